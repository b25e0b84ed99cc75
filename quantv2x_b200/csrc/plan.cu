// qv2x_plan: a fixed sequence of quantized layers over numbered activation buffers -- the whole
// BaseBEVBackbone + DownsampleConv forward of one modality (reference quant_block.py:280-303, :567-586)
// issued from C++ as back-to-back kernel launches on one stream: no host round trips, no torch.cat
// (deblocks write straight into channel slices of the concat buffer), no NCHW<->NHWC copies.
#include <array>
#include <map>
#include <vector>

#include "host_common.h"

struct qv2x_plan {
    std::vector<qv2x_plan_step> steps;
    std::vector<int> buf_channels;
    // Rowsum slots.  A slot holds the per-pixel channel sums of ONE tensor version: either a channel group of
    // the plan input (producer -1) or the output of one step.  Buffers are reused (ping-pong) inside a stage,
    // so slots are keyed by the producing step, never by the buffer.
    struct Slot {
        int producer;   // step index, or -1 for the plan input
        int buf, cbase, width;
    };
    std::vector<Slot> slots;
    std::vector<int> step_out_slot;                // per step: slot its epilogue accumulates into, or -1
    std::vector<std::array<int, 3>> step_in_slot;  // per step: slot of each input group, or -1
};

namespace {
struct Shapes {
    std::vector<int> h, w;
};

int propagate(const qv2x_plan* P, int H, int W, Shapes* s) {
    const int nb = static_cast<int>(P->buf_channels.size());
    s->h.assign(nb, -1);
    s->w.assign(nb, -1);
    s->h[0] = H;
    s->w[0] = W;
    for (const auto& st : P->steps) {
        if (s->h[st.in_buf] < 0) return qv2x::set_error(QV2X_ERR_INVALID, "plan step reads buffer %d before it is written", st.in_buf);
        int ho, wo;
        int rc = qv2x_layer_out_shape(st.layer, s->h[st.in_buf], s->w[st.in_buf], &ho, &wo);
        if (rc) return rc;
        if (ho <= 0 || wo <= 0) return qv2x::set_error(QV2X_ERR_INVALID, "input %dx%d too small for the plan", H, W);
        if (s->h[st.out_buf] >= 0 && (s->h[st.out_buf] != ho || s->w[st.out_buf] != wo))
            return qv2x::set_error(QV2X_ERR_INVALID, "writers of buffer %d disagree on its extent (%dx%d vs %dx%d)",
                                   st.out_buf, s->h[st.out_buf], s->w[st.out_buf], ho, wo);
        s->h[st.out_buf] = ho;
        s->w[st.out_buf] = wo;
    }
    return 0;
}

size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
}  // namespace

extern "C" {

int qv2x_plan_create(const qv2x_plan_step* steps, int n_steps, const int* buf_channels, int n_bufs, qv2x_plan** out) {
    QV2X_REQUIRE(steps && buf_channels && out && n_steps > 0 && n_bufs >= 2, "qv2x_plan_create: bad argument");
    auto P = new qv2x_plan();
    P->steps.assign(steps, steps + n_steps);
    P->buf_channels.assign(buf_channels, buf_channels + n_bufs);
    for (int i = 0; i < n_steps; ++i) {
        const auto& st = steps[i];
        if (!st.layer || st.in_buf < 0 || st.in_buf >= n_bufs || st.out_buf <= 0 || st.out_buf >= n_bufs) {
            delete P;
            return qv2x::set_error(QV2X_ERR_INVALID, "plan step %d: bad layer / buffer id", i);
        }
    }
    // rowsum slots: every consumer that applies the weight zero-point correction needs the channel sums of the
    // tensor version it reads = the most recent writer of (in_buf, channel base) before it
    P->step_out_slot.assign(n_steps, -1);
    P->step_in_slot.assign(n_steps, std::array<int, 3>{-1, -1, -1});
    std::map<std::pair<int, int>, int> last_writer;     // (buffer, cbase) -> step
    std::map<std::pair<int, int>, int> input_slot;      // (0, cbase) -> slot
    for (int i = 0; i < n_steps; ++i) {
        const auto& st = steps[i];
        qv2x_layer_desc d;
        qv2x_layer_desc_get(st.layer, &d);
        if (qv2x_layer_needs_rowsum(st.layer)) {
            const int cg = d.cin / d.n_in_groups;
            for (int g = 0; g < d.n_in_groups; ++g) {
                const auto key = std::make_pair(st.in_buf, st.in_cbase + g * cg);
                auto lw = last_writer.find(key);
                if (lw == last_writer.end()) {
                    if (st.in_buf != 0) {
                        delete P;
                        return qv2x::set_error(QV2X_ERR_INVALID, "plan step %d: nothing wrote channels [%d,+%d) of buffer %d",
                                               i, key.second, cg, st.in_buf);
                    }
                    if (!input_slot.count(key)) {
                        input_slot[key] = static_cast<int>(P->slots.size());
                        P->slots.push_back({-1, 0, key.second, cg});
                    }
                    P->step_in_slot[i][g] = input_slot[key];
                } else {
                    const int prod = lw->second;
                    qv2x_layer_desc pd;
                    qv2x_layer_desc_get(steps[prod].layer, &pd);
                    if (pd.cout != cg) {
                        delete P;
                        return qv2x::set_error(QV2X_ERR_INVALID,
                                               "plan step %d sums %d channels at (%d,%d) but step %d wrote %d there", i, cg,
                                               key.first, key.second, prod, pd.cout);
                    }
                    if (P->step_out_slot[prod] < 0) {
                        P->step_out_slot[prod] = static_cast<int>(P->slots.size());
                        P->slots.push_back({prod, st.in_buf, key.second, cg});
                    }
                    P->step_in_slot[i][g] = P->step_out_slot[prod];
                }
            }
        }
        last_writer[std::make_pair(st.out_buf, st.out_cbase)] = i;
    }
    *out = P;
    return 0;
}

void qv2x_plan_destroy(qv2x_plan* P) { delete P; }

int qv2x_plan_out_shape(const qv2x_plan* P, int H, int W, int* ho, int* wo, int* channels) {
    QV2X_REQUIRE(P && ho && wo && channels, "qv2x_plan_out_shape: null argument");
    Shapes s;
    int rc = propagate(P, H, W, &s);
    if (rc) return rc;
    const int last = static_cast<int>(P->buf_channels.size()) - 1;
    *ho = s.h[last];
    *wo = s.w[last];
    *channels = P->buf_channels[last];
    return 0;
}

int qv2x_plan_workspace_bytes(const qv2x_plan* P, int n_img, int H, int W, size_t* bytes) {
    QV2X_REQUIRE(P && bytes && n_img > 0, "qv2x_plan_workspace_bytes: bad argument");
    Shapes s;
    int rc = propagate(P, H, W, &s);
    if (rc) return rc;
    size_t total = 0;
    const int nb = static_cast<int>(P->buf_channels.size());
    for (int b = 1; b < nb - 1; ++b)
        if (s.h[b] > 0) total += align256(static_cast<size_t>(n_img) * s.h[b] * s.w[b] * P->buf_channels[b]);
    for (const auto& sl : P->slots)
        total += align256(static_cast<size_t>(n_img) * s.h[sl.buf] * s.w[sl.buf] * sizeof(int32_t));
    *bytes = total + 256;
    return 0;
}

int qv2x_plan_forward(const qv2x_plan* P, int n_img, int H, int W, const uint8_t* d_in, uint8_t* d_out,
                      void* d_workspace, size_t workspace_bytes, int dump_step, int32_t* d_acc_dump, void* stream_) {
    return qv2x_plan_forward_rs(P, n_img, H, W, d_in, nullptr, d_out, d_workspace, workspace_bytes, dump_step,
                                d_acc_dump, stream_);
}

int qv2x_plan_forward_rs(const qv2x_plan* P, int n_img, int H, int W, const uint8_t* d_in, const int32_t* d_in_rowsum,
                         uint8_t* d_out, void* d_workspace, size_t workspace_bytes, int dump_step, int32_t* d_acc_dump,
                         void* stream_) {
    QV2X_REQUIRE(P && d_in && d_out && d_workspace, "qv2x_plan_forward: null argument");
    size_t need = 0;
    int rc = qv2x_plan_workspace_bytes(P, n_img, H, W, &need);
    if (rc) return rc;
    QV2X_REQUIRE(workspace_bytes >= need, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Shapes s;
    propagate(P, H, W, &s);
    const int nb = static_cast<int>(P->buf_channels.size());
    std::vector<uint8_t*> buf(nb, nullptr);
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
    buf[0] = const_cast<uint8_t*>(d_in);
    buf[nb - 1] = d_out;
    for (int b = 1; b < nb - 1; ++b)
        if (s.h[b] > 0) {
            buf[b] = ws;
            ws += align256(static_cast<size_t>(n_img) * s.h[b] * s.w[b] * P->buf_channels[b]);
        }
    std::vector<int32_t*> slot(P->slots.size(), nullptr);
    uint8_t* rs_begin = ws;
    for (size_t k = 0; k < P->slots.size(); ++k) {
        const int b = P->slots[k].buf;
        slot[k] = reinterpret_cast<int32_t*>(ws);
        ws += align256(static_cast<size_t>(n_img) * s.h[b] * s.w[b] * sizeof(int32_t));
    }
    if (ws > rs_begin) QV2X_CUDA_OK(cudaMemsetAsync(rs_begin, 0, static_cast<size_t>(ws - rs_begin), stream));
    for (size_t k = 0; k < P->slots.size(); ++k) {
        if (P->slots[k].producer >= 0) continue;
        if (d_in_rowsum != nullptr && P->slots[k].cbase == 0 && P->slots[k].width == P->buf_channels[0]) {
            slot[k] = const_cast<int32_t*>(d_in_rowsum);      // supplied by the producer of the input
            continue;
        }
        rc = qv2x_rowsum_u8(d_in, static_cast<long long>(n_img) * H * W, P->buf_channels[0], P->slots[k].cbase,
                            P->slots[k].width, slot[k], stream);
        if (rc) return rc;
    }
    for (size_t i = 0; i < P->steps.size(); ++i) {
        const auto& st = P->steps[i];
        const int32_t* rs_in[3] = {nullptr, nullptr, nullptr};
        for (int g = 0; g < 3; ++g)
            if (P->step_in_slot[i][g] >= 0) rs_in[g] = slot[P->step_in_slot[i][g]];
        int32_t* rs_out = P->step_out_slot[i] >= 0 ? slot[P->step_out_slot[i]] : nullptr;
        rc = qv2x_layer_forward(st.layer, n_img, s.h[st.in_buf], s.w[st.in_buf], buf[st.in_buf],
                                P->buf_channels[st.in_buf], st.in_cbase, rs_in, buf[st.out_buf],
                                P->buf_channels[st.out_buf], st.out_cbase, rs_out,
                                (static_cast<int>(i) == dump_step) ? d_acc_dump : nullptr, stream);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
