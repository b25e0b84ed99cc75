// qv2x_postprocess: detection post-processing on the GPU:
//   head maps -> sigmoid, max over the classes, score threshold -> box decode against the anchors -> BEV corners ->
//   rotated NMS (top-1000 by score, greedy, IoU of convex quadrilaterals) -> range mask.
// Reference: VoxelPostprocessor3Heads.post_process (opencood/data_utils/post_processor/voxel_postprocessor_3heads.py:
// 318-477), generate_anchor_box (:63-127), delta_to_boxes3d (:581-635), boxes_to_corners_3d
// (opencood/utils/box_utils_mc.py:200-246), nms_rotated (box_utils_mc.py:665-710).  The reference runs this on the
// CPU inside its timed region, with shapely for the polygon IoU; here the arithmetic is float64 and follows
// oracle/postprocess_oracle.py operation by operation (Sutherland-Hodgman clipping + shoelace area).
//
// Four small kernels (the work is a few thousand candidates at most):
//   select : one thread per anchor; survivors are appended (atomic slot) with their decoded box and corners
//   rank   : rank by counting -- order by score descending, ties by larger anchor index first
//   iou    : suppression bit matrix of the top candidates (pair (i, j), i before j, IoU > threshold)
//   greedy : one warp walks the candidates in order and ORs the rows of the kept ones; range mask; compaction
#include <algorithm>
#include <cmath>
#include <vector>

#include "host_common.h"

namespace qv2x {

constexpr int kPPMaxClasses = 4, kPPMaxRot = 4, kPPTop = 1024, kPPWords = kPPTop / 32;

struct PPParams {
    int H, W, n_cls, n_rot;
    double x0[kPPMaxClasses], y0[kPPMaxClasses], xs[kPPMaxClasses], ys[kPPMaxClasses], z[kPPMaxClasses];
    double size[kPPMaxClasses][3];      // h, w, l
    double rot[kPPMaxRot];
    double score_thr;
    float logit_cut;                    // use_cut: logits <= logit_cut are below the score threshold for certain
    int use_cut;
    float nms_thr;
    double lo[2], hi[2];                // range mask on the BEV corners
    int cap, top;
};

struct PPCand {                          // one surviving anchor
    double score;
    int anchor, label;
    double box[7];
    double cx[4], cy[4];
};

__global__ void pp_select_kernel(const float* __restrict__ preds, long long hw, PPParams p, PPCand* __restrict__ cand,
                                 int* __restrict__ count) {
    const int A = p.n_cls * p.n_rot;
    const long long n = hw * A;
    for (long long a = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; a < n;
         a += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long pix = a / A;
        const int k = static_cast<int>(a - pix * A);
        if (p.use_cut) {
            // sigmoid is monotonic: logits at least 1e-6 below logit(threshold) cannot pass the float64 test below
            // (margin p (1 - p) 1e-6 >= 1e-10 against an evaluation error of ~1e-16), so the float64 exponentials
            // are only spent on the few anchors near or above the threshold
            bool any = false;
            for (int j = 0; j < p.n_cls; ++j)
                any = any || __ldg(preds + static_cast<long long>(k * p.n_cls + j) * hw + pix) > p.logit_cut;
            if (!any) continue;
        }
        double best = -1.0;
        int arg = 0;
        for (int j = 0; j < p.n_cls; ++j) {
            const double x = static_cast<double>(__ldg(preds + static_cast<long long>(k * p.n_cls + j) * hw + pix));
            const double pr = 1.0 / (1.0 + exp(-x));
            if (pr > best) best = pr, arg = j;           // first maximum, as numpy argmax
        }
        if (!(best > p.score_thr)) continue;
        const int slot = atomicAdd(count, 1);
        if (slot >= p.cap) continue;
        const int c = k / p.n_rot, r = k - c * p.n_rot;
        const int yi = static_cast<int>(pix / p.W), xi = static_cast<int>(pix - static_cast<long long>(yi) * p.W);
        const double an[7] = {p.x0[c] + xi * p.xs[c], p.y0[c] + yi * p.ys[c], p.z[c],
                              p.size[c][0], p.size[c][1], p.size[c][2], p.rot[r]};
        double d[7];
        const float* reg = preds + static_cast<long long>(p.n_cls * A + k * 7) * hw + pix;
        for (int t = 0; t < 7; ++t) d[t] = static_cast<double>(__ldg(reg + static_cast<long long>(t) * hw));
        PPCand o;
        o.score = best;
        o.anchor = static_cast<int>(a);
        o.label = arg + 1;
        const double diag = sqrt(an[4] * an[4] + an[5] * an[5]);
        o.box[0] = d[0] * diag + an[0];
        o.box[1] = d[1] * diag + an[1];
        o.box[2] = d[2] * an[3] + an[2];
        o.box[3] = exp(d[3]) * an[3];
        o.box[4] = exp(d[4]) * an[4];
        o.box[5] = exp(d[5]) * an[5];
        o.box[6] = d[6] + an[6];
        const double l = o.box[5], w = o.box[4], cs = cos(o.box[6]), sn = sin(o.box[6]);
        const double tx[4] = {0.5, 0.5, -0.5, -0.5}, ty[4] = {-0.5, 0.5, 0.5, -0.5};
        for (int q = 0; q < 4; ++q) {
            const double px = l * tx[q], py = w * ty[q];
            o.cx[q] = px * cs - py * sn + o.box[0];
            o.cy[q] = px * sn + py * cs + o.box[1];
        }
        cand[slot] = o;
    }
}

// order[rank] = candidate index, for rank < top (score descending; ties: larger anchor index first).
// One warp per candidate: the lanes share the count over the other candidates.
__global__ void pp_rank_kernel(const PPCand* __restrict__ cand, const int* __restrict__ count, int cap, int top,
                               int* __restrict__ order, int* __restrict__ n_top) {
    const int K = min(*count, cap);
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < K; i += nwarps) {
        const double s = cand[i].score;
        const int a = cand[i].anchor;
        int rank = 0;
        for (int j = lane; j < K; j += 32) {
            const double sj = cand[j].score;
            rank += (sj > s) || (sj == s && cand[j].anchor > a);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, off);
        if (lane == 0 && rank < top) order[rank] = i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_top = min(K, top);
}

struct Poly {
    double x[10], y[10];
    int n;
};
__device__ double poly_signed2(const double* x, const double* y, int n) {      // 2 * signed area (shoelace)
    double s1 = 0.0, s2 = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        s1 += x[i] * y[j];
        s2 += y[i] * x[j];
    }
    return s1 - s2;
}
// IoU of two convex quadrilaterals: Sutherland-Hodgman clip of p by q (q made counter-clockwise) + shoelace.
__device__ double quad_iou(const double* px, const double* py, const double* qx_in, const double* qy_in) {
    const double ap = 0.5 * fabs(poly_signed2(px, py, 4)), aq = 0.5 * fabs(poly_signed2(qx_in, qy_in, 4));
    double qx[4], qy[4];
    const bool rev = poly_signed2(qx_in, qy_in, 4) < 0;
    for (int i = 0; i < 4; ++i) qx[i] = qx_in[rev ? 3 - i : i], qy[i] = qy_in[rev ? 3 - i : i];
    Poly out;
    out.n = 4;
    for (int i = 0; i < 4; ++i) out.x[i] = px[i], out.y[i] = py[i];
    for (int e = 0; e < 4 && out.n > 0; ++e) {
        const double ax = qx[e], ay = qy[e], bx = qx[(e + 1) & 3], by = qy[(e + 1) & 3];
        Poly in = out;
        out.n = 0;
        double sx = in.x[in.n - 1], sy = in.y[in.n - 1];
        bool s_in = (bx - ax) * (sy - ay) - (by - ay) * (sx - ax) >= 0;
        for (int i = 0; i < in.n; ++i) {
            const double ex = in.x[i], ey = in.y[i];
            const bool e_in = (bx - ax) * (ey - ay) - (by - ay) * (ex - ax) >= 0;
            if (e_in != s_in) {           // the edge s -> e crosses the clip line
                const double d1x = bx - ax, d1y = by - ay, d2x = ex - sx, d2y = ey - sy;
                const double den = d1x * d2y - d1y * d2x;
                const double t = ((sx - ax) * d2y - (sy - ay) * d2x) / den;
                out.x[out.n] = ax + t * d1x;
                out.y[out.n] = ay + t * d1y;
                ++out.n;
            }
            if (e_in) {
                out.x[out.n] = ex;
                out.y[out.n] = ey;
                ++out.n;
            }
            sx = ex, sy = ey, s_in = e_in;
        }
    }
    const double inter = (out.n >= 3) ? 0.5 * fabs(poly_signed2(out.x, out.y, out.n)) : 0.0;
    const double uni = ap + aq - inter;
    return uni > 0 ? inter / uni : 0.0;
}

// sup[i][w] bit b: candidate order[32 w + b] (later in the order) overlaps candidate order[i] by more than the threshold
__global__ void pp_iou_kernel(const PPCand* __restrict__ cand, const int* __restrict__ order,
                              const int* __restrict__ n_top, float nms_thr, unsigned* __restrict__ sup) {
    const int K = *n_top;
    const int i = blockIdx.x;
    if (i >= K) return;
    const PPCand& ci = cand[order[i]];
    const double ix0 = fmin(fmin(ci.cx[0], ci.cx[1]), fmin(ci.cx[2], ci.cx[3]));
    const double ix1 = fmax(fmax(ci.cx[0], ci.cx[1]), fmax(ci.cx[2], ci.cx[3]));
    const double iy0 = fmin(fmin(ci.cy[0], ci.cy[1]), fmin(ci.cy[2], ci.cy[3]));
    const double iy1 = fmax(fmax(ci.cy[0], ci.cy[1]), fmax(ci.cy[2], ci.cy[3]));
    // words beyond the last candidate are never read by the greedy pass
    for (int w = threadIdx.x >> 5; w < (K + 31) / 32; w += blockDim.x >> 5) {
        const int j = 32 * w + (threadIdx.x & 31);
        bool bit = false;
        if (j > i && j < K) {
            const PPCand& cj = cand[order[j]];
            // quadrilaterals whose bounding boxes are disjoint have an empty clip (IoU 0, never above a positive
            // threshold): only overlapping pairs pay for the float64 clipping
            bool apart = false;
            if (nms_thr > 1e-6f) {
                const double jx0 = fmin(fmin(cj.cx[0], cj.cx[1]), fmin(cj.cx[2], cj.cx[3]));
                const double jx1 = fmax(fmax(cj.cx[0], cj.cx[1]), fmax(cj.cx[2], cj.cx[3]));
                const double jy0 = fmin(fmin(cj.cy[0], cj.cy[1]), fmin(cj.cy[2], cj.cy[3]));
                const double jy1 = fmax(fmax(cj.cy[0], cj.cy[1]), fmax(cj.cy[2], cj.cy[3]));
                apart = jx0 > ix1 || jx1 < ix0 || jy0 > iy1 || jy1 < iy0;
            }
            if (!apart) bit = static_cast<float>(quad_iou(ci.cx, ci.cy, cj.cx, cj.cy)) > nms_thr;
        }
        const unsigned word = __ballot_sync(0xffffffffu, bit);
        if ((threadIdx.x & 31) == 0) sup[i * kPPWords + w] = word;
    }
}

// One warp, no shared memory (it runs beside the persistent conv kernels of the next frame).  Phase 1 walks the
// candidates in order and ORs the suppression rows of the kept ones into the removed mask (lane w owns word w); the
// rows of the next 16 candidates are requested together, so the walk pays one memory latency per 16 candidates
// instead of one (plus two dependent ones) per candidate.  Phase 2 applies the range mask to 32 kept candidates at
// a time, one per lane, and compacts them in pick order.
__global__ void pp_greedy_kernel(const PPCand* __restrict__ cand, const int* __restrict__ order,
                                 const int* __restrict__ n_top, const unsigned* __restrict__ sup, PPParams p,
                                 double* __restrict__ out_corners, double* __restrict__ out_scores,
                                 int* __restrict__ out_labels, double* __restrict__ out_boxes, int* __restrict__ out_n) {
    const int lane = threadIdx.x;
    const int K = *n_top;
    const int nwords = (K + 31) / 32;
    unsigned removed = 0, kept = 0;       // lane w owns word w of both masks
    // this lane's candidates of phase 2 (rank 32 w + lane): their indices are requested before the walk starts
    int ord[kPPWords];
#pragma unroll
    for (int w = 0; w < kPPWords; ++w) ord[w] = (32 * w + lane < K) ? __ldg(order + 32 * w + lane) : 0;
    constexpr int kAhead = 16;
    auto load_rows = [&](int i0, unsigned (&row)[kAhead]) {
#pragma unroll
        for (int u = 0; u < kAhead; ++u)
            row[u] = (i0 + u < K && lane < nwords) ? __ldg(sup + (i0 + u) * kPPWords + lane) : 0u;
    };
    unsigned row[kAhead], nxt[kAhead];
    load_rows(0, row);
    for (int i0 = 0; i0 < K; i0 += kAhead) {
        load_rows(i0 + kAhead, nxt);       // in flight while this batch is walked
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
            const int i = i0 + u;
            const unsigned word = __shfl_sync(0xffffffffu, removed, i >> 5);
            const bool take = i < K && !((word >> (i & 31)) & 1u);
            if (take) {
                removed |= row[u];
                if (lane == (i >> 5)) kept |= 1u << (i & 31);
            }
        }
#pragma unroll
        for (int u = 0; u < kAhead; ++u) row[u] = nxt[u];
    }
    int n_out = 0;
#pragma unroll
    for (int w = 0; w < kPPWords; ++w) {
        if (w >= nwords) break;
        const unsigned kw = __shfl_sync(0xffffffffu, kept, w);
        if (kw == 0u) continue;
        const bool mine = (kw >> lane) & 1u;
        const PPCand* c = cand + ord[w];
        // every value of the candidate is requested before the first one is looked at (one memory latency per word)
        double cx[4], cy[4], box[7], score = 0.0;
        int label = 0;
        if (mine) {
#pragma unroll
            for (int q = 0; q < 4; ++q) cx[q] = c->cx[q], cy[q] = c->cy[q];
#pragma unroll
            for (int t = 0; t < 7; ++t) box[t] = c->box[t];
            score = c->score;
            label = c->label;
        }
        bool inside = mine;
        if (mine) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                inside = inside & (cx[q] >= p.lo[0]) & (cx[q] <= p.hi[0]) & (cy[q] >= p.lo[1]) & (cy[q] <= p.hi[1]);
        }
        const unsigned ok = __ballot_sync(0xffffffffu, inside);
        if (inside) {
            const int slot = n_out + __popc(ok & ((1u << lane) - 1u));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                out_corners[(slot * 4 + q) * 2 + 0] = cx[q];
                out_corners[(slot * 4 + q) * 2 + 1] = cy[q];
            }
#pragma unroll
            for (int t = 0; t < 7; ++t) out_boxes[slot * 7 + t] = box[t];
            out_scores[slot] = score;
            out_labels[slot] = label;
        }
        n_out += __popc(ok);
    }
    if (lane == 0) *out_n = n_out;
}

}  // namespace qv2x

using namespace qv2x;

struct qv2x_postprocess {
    PPParams p;
    PPCand* d_cand = nullptr;
    int *d_count = nullptr, *d_order = nullptr, *d_ntop = nullptr;
    unsigned* d_sup = nullptr;
};

extern "C" {

int qv2x_postprocess_create(const qv2x_postprocess_desc* d, qv2x_postprocess** out) {
    QV2X_REQUIRE(d && out, "qv2x_postprocess_create: null argument");
    QV2X_CHECK_SIZE(d, qv2x_postprocess_desc);
    QV2X_REQUIRE(d->n_classes >= 1 && d->n_classes <= kPPMaxClasses && d->n_rotations >= 1 &&
                     d->n_rotations <= kPPMaxRot, "1..%d classes, 1..%d rotations", kPPMaxClasses, kPPMaxRot);
    QV2X_REQUIRE(d->H > 0 && d->W > 0 && d->max_candidates >= 1 && d->top >= 1 && d->top <= kPPTop,
                 "bad map size / candidate caps (top <= %d)", kPPTop);
    auto h = new qv2x_postprocess();
    PPParams& p = h->p;
    p.H = d->H, p.W = d->W, p.n_cls = d->n_classes, p.n_rot = d->n_rotations;
    for (int c = 0; c < d->n_classes; ++c) {
        p.x0[c] = d->anchor_x0[c], p.y0[c] = d->anchor_y0[c], p.xs[c] = d->anchor_dx[c], p.ys[c] = d->anchor_dy[c];
        p.z[c] = d->anchor_z[c];
        for (int t = 0; t < 3; ++t) p.size[c][t] = d->anchor_hwl[c][t];
    }
    for (int r = 0; r < d->n_rotations; ++r) p.rot[r] = d->anchor_rot[r];
    p.score_thr = d->score_threshold;
    p.use_cut = (d->score_threshold > 1e-4 && d->score_threshold < 1.0 - 1e-4) ? 1 : 0;
    if (p.use_cut) {
        // the largest float that is still >= 1e-6 below logit(threshold)
        const double lc = std::log(d->score_threshold / (1.0 - d->score_threshold)) - 1e-6;
        float f = static_cast<float>(lc);
        if (static_cast<double>(f) > lc) f = std::nextafter(f, -INFINITY);
        p.logit_cut = f;
    }
    p.nms_thr = d->nms_threshold;
    p.lo[0] = d->range_lo[0], p.lo[1] = d->range_lo[1], p.hi[0] = d->range_hi[0], p.hi[1] = d->range_hi[1];
    p.cap = d->max_candidates, p.top = d->top;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&h->d_cand), sizeof(PPCand) * p.cap);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->d_count), sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->d_ntop), sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->d_order), sizeof(int) * kPPTop);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->d_sup), sizeof(unsigned) * kPPTop * kPPWords);
    if (e != cudaSuccess) {
        qv2x_postprocess_destroy(h);
        return set_error(QV2X_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    *out = h;
    return 0;
}

void qv2x_postprocess_destroy(qv2x_postprocess* h) {
    if (!h) return;
    cudaFree(h->d_cand);
    cudaFree(h->d_count);
    cudaFree(h->d_ntop);
    cudaFree(h->d_order);
    cudaFree(h->d_sup);
    delete h;
}

int qv2x_postprocess_forward(const qv2x_postprocess* h, const float* d_preds, double* d_corners, double* d_scores,
                             int* d_labels, double* d_boxes, int* d_n_out, int* d_n_candidates, void* stream_) {
    QV2X_REQUIRE(h && d_preds && d_corners && d_scores && d_labels && d_boxes && d_n_out,
                 "qv2x_postprocess_forward: null argument");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const PPParams& p = h->p;
    const long long hw = static_cast<long long>(p.H) * p.W;
    QV2X_CUDA_OK(cudaMemsetAsync(h->d_count, 0, sizeof(int), stream));
    const long long n = hw * p.n_cls * p.n_rot;
    pp_select_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, num_sms() * 8LL)), 256, 0, stream>>>(
        d_preds, hw, p, h->d_cand, h->d_count);
    pp_rank_kernel<<<std::max(1, std::min((p.cap + 7) / 8, 2 * num_sms())), 256, 0, stream>>>(
        h->d_cand, h->d_count, p.cap, p.top, h->d_order, h->d_ntop);
    pp_iou_kernel<<<p.top, 256, 0, stream>>>(h->d_cand, h->d_order, h->d_ntop, p.nms_thr, h->d_sup);
    pp_greedy_kernel<<<1, 32, 0, stream>>>(h->d_cand, h->d_order, h->d_ntop, h->d_sup, p, d_corners, d_scores,
                                           d_labels, d_boxes, d_n_out);
    g_launch_count.fetch_add(4);
    if (d_n_candidates)
        QV2X_CUDA_OK(cudaMemcpyAsync(d_n_candidates, h->d_count, sizeof(int), cudaMemcpyDeviceToDevice, stream));
    QV2X_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
