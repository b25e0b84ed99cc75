// qv2x_layer: one quantized conv / transposed-conv layer (C ABI in include/qv2x.h).
// Host side: weight packing, tile/launch selection, tensor maps.  Device side: igemm.cuh.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "layer_launch.h"



using namespace qv2x;

struct qv2x_layer {
    qv2x_layer_desc d;
    int groups;        // accumulator groups in the kernel
    int n_total;       // GEMM columns
    int k_total;       // bytes per B row
    int block_n, bk;
    bool b_signed, use_zp;
    uint8_t* d_w = nullptr;
    float* d_cscale = nullptr;
    float* d_bias = nullptr;
    int32_t* d_zpw = nullptr;
    float gscale[3] = {1.f, 1.f, 1.f};
    // fixed-point requantization (epilogue_fixed.cuh) when every column's parameters are representable
    bool fixed_ok = false;
    int32_t* d_mul[3] = {nullptr, nullptr, nullptr};
    int2* d_cadd = nullptr;
    int32_t* d_shr = nullptr;
    // host copies of the same parameters: layers of up to 256 columns pass them through the kernel-parameter bank
    std::vector<int32_t> h_mul[3], h_shr, h_zpw;
    std::vector<int2> h_cadd;
};

namespace {
// One output column of the fixed-point requantizer: M_g = rint(r_g * 2^sh), C = rint(bq * 2^(sh - c_off)) +
// 2^(sh - c_off - 1), shift of the high word = sh - sh_min.  Returns false when the column does not fit
// (the layer then keeps the fp32 epilogue).  oracle/int_oracle.py::fixed_column is the same arithmetic.
bool fixed_column(const double* r, int ng, double bq, int sh_min, int sh_max, int c_off, int32_t* M, int2* cadd,
                  int32_t* shr) {
    double rmax = 0.0;
    for (int g = 0; g < ng; ++g) {
        if (!(r[g] >= 0.0) || !std::isfinite(r[g])) return false;
        rmax = std::max(rmax, r[g]);
    }
    if (!(rmax > 0.0) || !std::isfinite(bq) || std::fabs(bq) >= 32768.0) return false;
    int e = 0;
    std::frexp(rmax, &e);
    int sh = 31 - e;
    if (sh < sh_min) return false;
    if (sh > sh_max) sh = sh_max;
    long long m[3] = {0, 0, 0};
    for (;;) {
        bool ok = true;
        for (int g = 0; g < ng; ++g) {
            m[g] = std::llrint(std::ldexp(r[g], sh));
            if (m[g] >= (1LL << 31)) ok = false;
        }
        if (ok) break;
        if (--sh < sh_min) return false;
    }
    const int cs = sh - c_off;
    const long long c = std::llrint(std::ldexp(bq, cs)) + (1LL << (cs - 1));
    for (int g = 0; g < ng; ++g) M[g] = static_cast<int32_t>(m[g]);
    cadd->x = static_cast<int32_t>(static_cast<uint32_t>(c & 0xffffffffLL));
    cadd->y = static_cast<int32_t>(c >> 32);
    *shr = sh - sh_min;
    return true;
}
}  // namespace

extern "C" {

int qv2x_layer_create(const qv2x_layer_desc* desc, const uint8_t* w_int, const float* w_delta,
                      const float* w_zero_point, const float* bias, qv2x_layer** out) {
    QV2X_REQUIRE(desc && w_int && w_delta && w_zero_point && out, "qv2x_layer_create: null argument");
    QV2X_CHECK_SIZE(desc, qv2x_layer_desc);
    const qv2x_layer_desc& d = *desc;
    QV2X_REQUIRE(d.kind == 0 || d.kind == 1, "kind must be 0 (conv) or 1 (transposed conv)");
    QV2X_REQUIRE(d.w_bits >= 2 && d.w_bits <= 8 && d.out_bits >= 2 && d.out_bits <= 8, "bit widths must be 2..8");
    QV2X_REQUIRE(d.n_in_groups == 1 || d.n_in_groups == 3, "n_in_groups must be 1 or 3");
    QV2X_REQUIRE(d.out_delta > 0.f, "out_delta must be positive");
    const int wg = d.groups > 1 ? d.groups : 1;       // nn.Conv2d(groups=...)
    QV2X_REQUIRE(wg == 1 || (d.kind == 0 && d.n_in_groups == 1 && d.cin % wg == 0 && d.cout % wg == 0),
                 "grouped weights: conv only, one input scale, channels divisible by groups");
    auto L = new qv2x_layer();
    L->d = d;
    std::vector<uint8_t> wpack;
    std::vector<float> cscale, biasv;
    std::vector<int32_t> zpw;
    std::vector<int32_t> fx_mul[3], fx_shr;
    std::vector<int2> fx_cadd;
    bool fx_ok = false;

    if (d.kind == 0) {
        QV2X_REQUIRE((d.ksize == 1 || d.ksize == 3) && (d.stride == 1 || d.stride == 2), "conv: ksize 1|3, stride 1|2");
        QV2X_REQUIRE(d.cin % d.n_in_groups == 0, "cin must split evenly over the input groups");
        const int cg = d.cin / d.n_in_groups;
        QV2X_REQUIRE(cg % 64 == 0, "input channels per group must be a multiple of 64 (got %d)", cg);
        QV2X_REQUIRE(d.cout % 64 == 0, "cout must be a multiple of 64 (got %d)", d.cout);
        L->groups = d.n_in_groups;
        L->bk = (cg % 128 == 0) ? 128 : 64;
        L->n_total = d.cout;
        if (L->groups == 1 && d.cout % 256 == 0) L->block_n = 256;
        else if (d.cout % 128 == 0) L->block_n = 128;
        else L->block_n = 64;
        const int taps = d.ksize * d.ksize;
        L->k_total = taps * d.cin;
        L->b_signed = d.w_bits <= 7;   // (w - zp) in [-127, 127] fits a signed byte: no zero-point correction
        L->use_zp = !L->b_signed;
        wpack.resize(static_cast<size_t>(d.cout) * L->k_total);
        cscale.resize(d.cout);
        biasv.resize(d.cout);
        zpw.resize(d.cout);
        // Grouped weights ([cout][cin/groups][k][k], ResNeXt convs of the pyramid backbone) run as the dense GEMM of
        // the block-diagonal matrix: an off-group entry is the channel's zero-point, i.e. a real weight of exactly 0,
        // so every integer accumulator equals the grouped convolution's.  (Groups of 4-16 channels are far below
        // the 32-byte K step of the tensor pipe; the dense form costs `groups` x the MACs at full tensor rate.)
        const int cin_g = d.cin / wg, cout_g = d.cout / wg;
        for (int co = 0; co < d.cout; ++co) {
            const int zp = static_cast<int>(w_zero_point[co]);
            zpw[co] = zp;
            // G=1: y = float(acc) * fl(in_delta * w_delta[c]);  G=3: y = (sum_g in_delta[g]*acc_g) * w_delta[c]
            cscale[co] = (L->groups == 1) ? d.in_delta[0] * w_delta[co] : w_delta[co];
            biasv[co] = bias ? bias[co] : 0.f;
            const int ci0 = (co / cout_g) * cin_g;
            for (int ci = 0; ci < d.cin; ++ci)
                for (int t = 0; t < taps; ++t) {
                    const bool in_group = (ci >= ci0 && ci < ci0 + cin_g);
                    const int w = in_group ? w_int[(static_cast<size_t>(co) * cin_g + (ci - ci0)) * taps + t] : zp;
                    const int v = L->b_signed ? (w - zp) : w;
                    wpack[static_cast<size_t>(co) * L->k_total + static_cast<size_t>(t) * d.cin + ci] =
                        static_cast<uint8_t>(v & 0xff);
                }
        }
        for (int g = 0; g < 3; ++g) L->gscale[g] = (L->groups == 1) ? 1.f : d.in_delta[g];
        fx_ok = (d.out_bits == 8 && d.out_zero_point == 0.f && d.relu);
        for (int g = 0; g < L->groups; ++g) fx_mul[g].resize(d.cout);
        fx_cadd.resize(d.cout);
        fx_shr.resize(d.cout);
        for (int co = 0; co < d.cout && fx_ok; ++co) {
            double r[3];
            for (int g = 0; g < L->groups; ++g)
                r[g] = static_cast<double>(L->gscale[g]) * static_cast<double>(cscale[co]) / static_cast<double>(d.out_delta);
            int32_t m[3];
            fx_ok = fixed_column(r, L->groups, static_cast<double>(biasv[co]) / static_cast<double>(d.out_delta), 32, 47, 0,
                                 m, &fx_cadd[co], &fx_shr[co]);
            for (int g = 0; g < L->groups && fx_ok; ++g) fx_mul[g][co] = m[g];
        }
    } else {
        QV2X_REQUIRE(d.ksize == d.stride && (d.stride == 1 || d.stride == 2 || d.stride == 4),
                     "transposed conv: kernel == stride in {1,2,4}");
        QV2X_REQUIRE(d.n_in_groups == 1 && d.pad == 0, "transposed conv: single input group, padding 0");
        QV2X_REQUIRE(d.cin % 64 == 0 && d.cout % 64 == 0, "channels must be multiples of 64");
        // Weight scale varies along the reduction (cin) axis (reference quant_layer.py:325-335 on a
        // [cin, cout, k, k] tensor), so the real-valued weight w_hat = (W - zp[ci]) * delta[ci] is carried
        // as a 24-bit fixed-point number per output column, split into three signed byte digits.
        const int s = d.stride, sub = s * s;
        L->groups = 3;
        L->bk = (d.cin % 128 == 0) ? 128 : 64;
        L->block_n = (d.cout % 128 == 0) ? 128 : 64;
        L->n_total = sub * d.cout;
        L->k_total = d.cin;
        L->b_signed = true;
        L->use_zp = false;
        wpack.assign(static_cast<size_t>(3) * L->n_total * d.cin, 0);
        cscale.resize(L->n_total);
        biasv.resize(L->n_total);
        std::vector<float> what(d.cin);
        for (int sp = 0; sp < sub; ++sp)
            for (int co = 0; co < d.cout; ++co) {
                const int n = sp * d.cout + co;
                double mx = 0.0;
                for (int ci = 0; ci < d.cin; ++ci) {
                    const float wq = static_cast<float>(w_int[(static_cast<size_t>(ci) * d.cout + co) * sub + sp]);
                    what[ci] = (wq - w_zero_point[ci]) * w_delta[ci];   // fp32, as the reference dequantizes
                    mx = std::max(mx, std::fabs(static_cast<double>(what[ci])));
                }
                const double sc = mx > 0.0 ? mx / kDigitMax : 1.0;
                for (int ci = 0; ci < d.cin; ++ci) {
                    const long long m = static_cast<long long>(std::nearbyint(static_cast<double>(what[ci]) / sc));
                    int8_t dg[3];
                    split_digits(m, dg);
                    for (int g = 0; g < 3; ++g)
                        wpack[(static_cast<size_t>(g) * L->n_total + n) * d.cin + ci] = static_cast<uint8_t>(dg[g]);
                }
                cscale[n] = static_cast<float>(static_cast<double>(d.in_delta[0]) * sc);
                biasv[n] = bias ? bias[co] : 0.f;
            }
        L->gscale[0] = 65536.f;
        L->gscale[1] = 256.f;
        L->gscale[2] = 1.f;
        fx_ok = (d.out_bits == 8 && d.out_zero_point == 0.f && d.relu && d.cin <= 256);
        fx_mul[0].resize(L->n_total);
        fx_cadd.resize(L->n_total);
        fx_shr.resize(L->n_total);
        for (int n = 0; n < L->n_total && fx_ok; ++n) {
            const double r = static_cast<double>(cscale[n]) / static_cast<double>(d.out_delta);
            fx_ok = fixed_column(&r, 1, static_cast<double>(biasv[n]) / static_cast<double>(d.out_delta), 48, 62, 16,
                                 &fx_mul[0][n], &fx_cadd[n], &fx_shr[n]);
        }
    }
    int rc = upload(&L->d_w, wpack.data(), wpack.size());
    if (rc == 0) rc = upload(&L->d_cscale, cscale.data(), cscale.size());
    if (rc == 0) rc = upload(&L->d_bias, biasv.data(), biasv.size());
    if (rc == 0 && L->use_zp) rc = upload(&L->d_zpw, zpw.data(), zpw.size());
    static const int env_fixed = getenv("QV2X_FIXED") ? atoi(getenv("QV2X_FIXED")) : 1;
    if (rc == 0 && fx_ok && env_fixed) {
        for (int g = 0; g < 3 && rc == 0; ++g)
            if (!fx_mul[g].empty()) rc = upload(&L->d_mul[g], fx_mul[g].data(), fx_mul[g].size());
        if (rc == 0) rc = upload(&L->d_cadd, fx_cadd.data(), fx_cadd.size());
        if (rc == 0) rc = upload(&L->d_shr, fx_shr.data(), fx_shr.size());
        L->fixed_ok = (rc == 0);
        for (int g = 0; g < 3; ++g) L->h_mul[g] = fx_mul[g];
        L->h_cadd = fx_cadd;
        L->h_shr = fx_shr;
        L->h_zpw = zpw;
    }
    if (rc != 0) {
        qv2x_layer_destroy(L);
        return rc;
    }
    *out = L;
    return 0;
}

void qv2x_layer_destroy(qv2x_layer* L) {
    if (!L) return;
    cudaFree(L->d_w);
    cudaFree(L->d_cscale);
    cudaFree(L->d_bias);
    cudaFree(L->d_zpw);
    for (int g = 0; g < 3; ++g) cudaFree(L->d_mul[g]);
    cudaFree(L->d_cadd);
    cudaFree(L->d_shr);
    delete L;
}

int qv2x_layer_needs_rowsum(const qv2x_layer* L) { return (L && L->use_zp) ? 1 : 0; }

int qv2x_layer_desc_get(const qv2x_layer* L, qv2x_layer_desc* out) {
    QV2X_REQUIRE(L && out, "qv2x_layer_desc_get: null argument");
    *out = L->d;
    return 0;
}

int qv2x_layer_out_shape(const qv2x_layer* L, int hi, int wi, int* ho, int* wo) {
    QV2X_REQUIRE(L && ho && wo, "qv2x_layer_out_shape: null argument");
    const qv2x_layer_desc& d = L->d;
    if (d.kind == 0) {
        *ho = (hi + 2 * d.pad - d.ksize) / d.stride + 1;
        *wo = (wi + 2 * d.pad - d.ksize) / d.stride + 1;
    } else {
        *ho = hi * d.stride;
        *wo = wi * d.stride;
    }
    return 0;
}

int qv2x_layer_forward(const qv2x_layer* L, int n_img, int hi, int wi, const uint8_t* d_x, int in_cstride,
                       int in_cbase, const int32_t* const* d_rowsum_in, uint8_t* d_y, int out_cstride, int out_cbase,
                       int32_t* d_rowsum_out, int32_t* d_acc_dump, void* stream_) {
    return qv2x_layer_forward_ex(L, n_img, hi, wi, d_x, in_cstride, in_cbase, d_rowsum_in, d_y, out_cstride, out_cbase,
                                 d_rowsum_out, d_acc_dump, nullptr, stream_);
}

int qv2x_layer_forward_ex(const qv2x_layer* L, int n_img, int hi, int wi, const uint8_t* d_x, int in_cstride,
                          int in_cbase, const int32_t* const* d_rowsum_in, uint8_t* d_y, int out_cstride,
                          int out_cbase, int32_t* d_rowsum_out, int32_t* d_acc_dump, const qv2x_layer_extra* ex,
                          void* stream_) {
    QV2X_REQUIRE(L && d_x, "qv2x_layer_forward: null argument");
    const bool f32_out = ex && ex->d_out_f32;
    const bool has_res = ex && (ex->d_res_u8 || ex->d_res_f32);
    QV2X_REQUIRE(d_y || f32_out, "qv2x_layer_forward: null output");
    if (ex) {
        QV2X_CHECK_SIZE(ex, qv2x_layer_extra);
        QV2X_REQUIRE(!(ex->d_res_u8 && ex->d_res_f32), "one shortcut at most");
        QV2X_REQUIRE(!ex->d_res_u8 || (ex->res_cstride % 4 == 0 && ex->res_cbase % 4 == 0 && ex->res_delta > 0.f &&
                                       (reinterpret_cast<uintptr_t>(ex->d_res_u8) & 3) == 0),
                     "code shortcut: 4-byte aligned, positive scale");
        QV2X_REQUIRE(!ex->d_res_f32 || (ex->res_cstride % 4 == 0 && ex->res_cbase % 4 == 0 &&
                                        (reinterpret_cast<uintptr_t>(ex->d_res_f32) & 15) == 0),
                     "FP32 shortcut: 16-byte aligned rows");
        QV2X_REQUIRE(!f32_out || (ex->out_f32_cstride % 4 == 0 && ex->out_f32_cstride >= L->d.cout &&
                                  (reinterpret_cast<uintptr_t>(ex->d_out_f32) & 15) == 0),
                     "FP32 output: 16-byte aligned rows of at least cout floats");
    }
    if (!d_y) d_y = reinterpret_cast<uint8_t*>(ex->d_out_f32);    // never written on the FP32-output path
    QV2X_REQUIRE(n_img > 0 && hi > 0 && wi > 0, "empty input (n_img=%d hi=%d wi=%d)", n_img, hi, wi);
    QV2X_REQUIRE(in_cstride % 16 == 0 && in_cbase % 16 == 0 && out_cstride % 16 == 0 && out_cbase % 16 == 0,
                 "channel strides / bases must be multiples of 16 bytes");
    QV2X_REQUIRE((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_y) & 15) == 0,
                 "activation pointers must be 16-byte aligned");
    QV2X_REQUIRE(!(f32_out && d_rowsum_out), "an FP32 output has no code sums");
    QV2X_REQUIRE(!L->use_zp || d_rowsum_in, "this layer needs d_rowsum_in (see qv2x_layer_needs_rowsum)");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const qv2x_layer_desc& d = L->d;
    int ho, wo;
    qv2x_layer_out_shape(L, hi, wi, &ho, &wo);
    QV2X_REQUIRE(ho > 0 && wo > 0, "input %dx%d too small for this layer", hi, wi);

    IgemmGeom g{};
    g.n_img = n_img;
    g.Hi = hi;
    g.Wi = wi;
    g.groups = L->groups;
    g.n_steps = 1;
    int out_h, out_w;
    if (d.kind == 0) {
        g.Ho = ho;
        g.Wo = wo;
        g.taps_w = d.ksize;
        g.taps = d.ksize * d.ksize;
        g.stride = d.stride;
        g.pad = d.pad;
        const int cg = d.cin / d.n_in_groups;
        g.cblocks = cg / L->bk;
        g.b_k_tap_stride = d.cin;
        for (int i = 0; i < L->groups; ++i) {
            g.a_c_base[i] = in_cbase + i * cg;
            g.b_row_base[i] = 0;
            g.b_k_base[i] = i * cg;
        }
        out_h = ho;
        out_w = wo;
    } else {
        g.Ho = hi;   // GEMM rows are INPUT pixels; each N tile scatters to one output sub-position
        g.Wo = wi;
        g.taps_w = 1;
        g.taps = 1;
        g.stride = 1;
        g.pad = 0;
        g.cblocks = d.cin / L->bk;
        g.b_k_tap_stride = 0;
        // B rows are stored digit-major (hi, mid, lo).  With at most 256 reduction terms 256 * acc_mid + acc_lo
        // fits an int32, so the digits are reduced in the order mid, lo, hi and combined as
        // fl32(65536 * hi + fl32(256 * mid + lo)) -- two conversions and one fma instead of three of each.
        const bool packed = (d.cin <= 256);
        const int order[3] = {packed ? 1 : 0, packed ? 2 : 1, packed ? 0 : 2};
        for (int i = 0; i < 3; ++i) {
            g.a_c_base[i] = in_cbase;
            g.b_row_base[i] = order[i] * L->n_total;
            g.b_k_base[i] = 0;
        }
        out_h = ho;
        out_w = wo;
    }
    // 3x3 stride-1 convs take the HALO mainloop: fixed 8 x 16 pixel tiles, the input window loaded once per tile
    static const int env_halo = getenv("QV2X_HALO") ? atoi(getenv("QV2X_HALO")) : 1;
    const bool halo = env_halo && d.kind == 0 && d.ksize == 3 && d.stride == 1;
    if (halo) {
        g.tw = kHaloTileW;
        g.th = kHaloTileH;
    } else {
        choose_tile_box(g.Ho, g.Wo, &g.tw, &g.th);
    }
    g.tiles_x = (g.Wo + g.tw - 1) / g.tw;
    g.tiles_y = (g.Ho + g.th - 1) / g.th;
    {   // the side warp's halo of per-pixel input sums must fit its shared-memory buffer
        const int hw = (g.tw - 1) * g.stride + g.taps_w, hh = (g.th - 1) * g.stride + g.taps / g.taps_w;
        QV2X_REQUIRE(hw * hh <= RequantEpilogue<1>::kHaloInts, "halo %d x %d exceeds the side buffer", hh, hw);
    }
    // Column-tile width: the widest tile has the best operand reuse, but small maps (deep stages, one agent per
    // GPU) would leave most SMs idle or pay a nearly empty second wave.  Cost model in SM cycles per CTA:
    // waves x (fixed per-tile overhead: pipeline fill + the last tile's epilogue tail, + MMA time K/32 x bn/2).
    int block_n = L->block_n;
    {
        const long long m_tiles = static_cast<long long>(n_img) * g.tiles_x * g.tiles_y;
        const int sub_cols = (d.kind == 0) ? L->n_total : d.cout;   // a tile must not straddle sub-positions
        const long long k_steps = static_cast<long long>(L->groups) * g.taps * g.cblocks * (L->bk / 32);
        long long best = -1;
        static const int env_bn = getenv("QV2X_BN") ? atoi(getenv("QV2X_BN")) : 0;      // experiments
        for (int bn = L->block_n; bn >= 64; bn >>= 1) {
            if (sub_cols % bn != 0) continue;
            if (env_bn && bn != env_bn && sub_cols % env_bn == 0 && env_bn <= L->block_n) continue;
            const long long tiles = m_tiles * (L->n_total / bn);
            const long long waves = (tiles + num_sms() - 1) / num_sms();
            const long long cost = waves * (4500 + k_steps * (bn / 2));
            if (best < 0 || cost < best) {
                best = cost;
                block_n = bn;
            }
        }
    }
    HaloPlan hp{};
    if (halo) {
        // tile width and weights-resident mode by the cost model of plan_halo (QV2X_HALO_BN / QV2X_HALO_RES override
        // it for experiments)
        static const int env_bn = getenv("QV2X_HALO_BN") ? atoi(getenv("QV2X_HALO_BN")) : 0;
        static const int env_res = getenv("QV2X_HALO_RES") ? atoi(getenv("QV2X_HALO_RES")) : -1;
        const long long m_tiles = static_cast<long long>(n_img) * g.tiles_x * g.tiles_y;
        hp.cost = -1;
        for (int bn = 256; bn >= 64; bn >>= 1) {
            if (L->n_total % bn != 0 || (bn == 256 && (L->groups != 1 || L->bk != 128))) continue;
            if (env_bn && bn != env_bn && L->n_total % env_bn == 0 && !(env_bn == 256 && (L->groups != 1 || L->bk != 128)))
                continue;
            for (int res = 1; res >= 0; --res) {
                if (env_res >= 0 && res != env_res) continue;
                const HaloPlan c = plan_halo(bn, L->bk, L->groups, g.cblocks, m_tiles, L->n_total, res == 1, res);
                if (c.cost < 0 || (res == 1 && !c.resident)) continue;
                if (hp.cost < 0 || c.cost < hp.cost) hp = c;
            }
        }
        if (hp.cost < 0) {   // a forced mode that does not fit: fall back to the streamed default
            hp = plan_halo(L->block_n == 256 && L->bk != 128 ? 128 : L->block_n, L->bk, L->groups, g.cblocks, m_tiles,
                           L->n_total, false, 0);
        }
        QV2X_REQUIRE(hp.cost >= 0, "no HALO configuration fits shared memory");
        block_n = hp.block_n;
    }
    g.block_n = block_n;
    g.n_tiles = L->n_total / block_n;
    g.idesc = make_idesc_i8(block_n, L->b_signed);

    CUtensorMap tmA, tmB;
    int rc = halo ? make_halo_tmap(&tmA, d_x, n_img, hi, wi, in_cstride, L->bk)
                  : make_act_tmap(&tmA, d_x, n_img, hi, wi, in_cstride, g.tw, g.th, g.stride, L->bk);
    if (rc) return rc;
    rc = make_weight_tmap(&tmB, L->d_w, (d.kind == 0 ? 1 : 3) * L->n_total, L->k_total, block_n, L->bk);
    if (rc) return rc;
    LayerLaunch ctx{block_n, L->bk, halo, tmA, tmB, g, hp, stream};

    auto fill = [&](auto& e) {
        e.up = (d.kind == 0) ? 1 : d.stride;
        e.cout_sub = d.cout;
        e.up_shift = (e.up == 4) ? 2 : (e.up == 2 ? 1 : 0);
        e.debug = g_debug_flags;
        e.fd_cout_sub = FastDiv(d.cout);
        e.Hout = out_h;
        e.Wout = out_w;
        e.out_cstride = out_cstride;
        e.out_cbase = out_cbase;
        e.relu = d.relu;
        e.qmax = static_cast<float>((1 << d.out_bits) - 1);
        e.delta_out = d.out_delta;
        e.zp_out = d.out_zero_point;
        e.rdelta = 1.0f / d.out_delta;
        for (int i = 0; i < 3; ++i) {
            e.gscale[i] = L->gscale[i];
            e.zpw[i] = (L->use_zp && i < L->groups) ? L->d_zpw : nullptr;
            e.rowsum_in[i] = (L->use_zp && i < L->groups) ? d_rowsum_in[i] : nullptr;
        }
        e.cscale = L->d_cscale;
        e.bias = L->d_bias;
        e.out = d_y;
        e.rowsum_out = d_rowsum_out;
        e.acc_dump = d_acc_dump;
        e.n_total = L->n_total;
        e.res_u8 = ex ? ex->d_res_u8 : nullptr;
        e.res_f32 = ex ? ex->d_res_f32 : nullptr;
        e.res_delta = ex ? ex->res_delta : 0.f;
        e.res_cstride = ex ? ex->res_cstride : 0;
        e.res_cbase = ex ? ex->res_cbase : 0;
        e.out_f32 = ex ? ex->d_out_f32 : nullptr;
        e.out_f32_cstride = ex ? ex->out_f32_cstride : 0;
    };
    // 8-bit / zero-point 0 / ReLU outputs without a shortcut requantize in 64-bit fixed point (epilogue_fixed.cuh)
    static const int env_fixed_g3 = getenv("QV2X_FIXED_G3") ? atoi(getenv("QV2X_FIXED_G3")) : 1;
    if (L->fixed_ok && !f32_out && !has_res && (env_fixed_g3 || d.kind == 1 || L->groups == 1)) {
        auto fill_fx = [&](auto& e) {
            e.up = (d.kind == 0) ? 1 : d.stride;
            e.cout_sub = d.cout;
            e.up_shift = (e.up == 4) ? 2 : (e.up == 2 ? 1 : 0);
            e.debug = g_debug_flags;
            e.fd_cout_sub = FastDiv(d.cout);
            e.Hout = out_h;
            e.Wout = out_w;
            e.out_cstride = out_cstride;
            e.out_cbase = out_cbase;
            for (int i = 0; i < 3; ++i) {
                e.mul[i] = L->d_mul[i];
                e.rowsum_in[i] = (L->use_zp && i < L->groups) ? d_rowsum_in[i] : nullptr;
            }
            e.cadd = L->d_cadd;
            e.shr = L->d_shr;
            e.zpw = L->use_zp ? L->d_zpw : nullptr;
            e.out = d_y;
            e.rowsum_out = d_rowsum_out;
            e.acc_dump = d_acc_dump;
            e.n_total = L->n_total;
        };
#define QV2X_FX(GG, DG, ZPP, DMP)                                                          \
    {                                                                                      \
        FixedEpilogue<GG, DG, ZPP, DMP> e{};                                               \
        fill_fx(e);                                                                        \
        return run_layer(ctx, e);                                                          \
    }
        const bool dump = (d_acc_dump != nullptr);
        static const int env_cpar = getenv("QV2X_CPAR") ? atoi(getenv("QV2X_CPAR")) : 1;
        // (constant-bank parameters pay off for one-group layers whose tile spans all output columns: two code copies,
        //  5 KB of constants; measured on s0 / s1 / s2 / shrinker, profiles/r2_sweep_epilogue_params.log)
        if (d.kind == 0 && !dump && L->n_total <= 256 && env_cpar && L->groups == 1 && g.n_tiles == 1) {
            auto fill_tab = [&](auto& e) {
                fill_fx(e);
                for (int n = 0; n < L->n_total; ++n) {
                    for (int q = 0; q < L->groups; ++q) e.tab.mul[q][n] = L->h_mul[q][n];
                    e.tab.cadd[n] = L->h_cadd[n];
                    e.tab.shr[n] = L->h_shr[n];
                    e.tab.zw[n] = L->use_zp ? L->h_zpw[n] : 0;
                }
            };
#define QV2X_FXC(GG, ZPP)                 \
    {                                     \
        FixedEpilogueC<GG, ZPP> e{};      \
        fill_tab(e);                      \
        return run_layer(ctx, e);         \
    }
            if (L->groups == 1) {
                if (L->use_zp) QV2X_FXC(1, true) else QV2X_FXC(1, false)
            }
            if (L->use_zp) QV2X_FXC(3, true) else QV2X_FXC(3, false)
#undef QV2X_FXC
        }
        if (d.kind == 1) {
            if (dump) QV2X_FX(3, true, false, true) else QV2X_FX(3, true, false, false)
        }
        if (L->groups == 1) {
            if (dump) {
                if (L->use_zp) QV2X_FX(1, false, true, true) else QV2X_FX(1, false, false, true)
            }
            if (L->use_zp) QV2X_FX(1, false, true, false) else QV2X_FX(1, false, false, false)
        }
        if (dump) {
            if (L->use_zp) QV2X_FX(3, false, true, true) else QV2X_FX(3, false, false, true)
        }
        if (L->use_zp) QV2X_FX(3, false, true, false) else QV2X_FX(3, false, false, false)
#undef QV2X_FX
    }
    // FP32 outputs live in the generic epilogue only; a shortcut has a saturating fast variant for one input group
    const bool sat8 = (d.out_bits == 8 && d.out_zero_point == 0.f && d.relu) && !f32_out;
    const bool digits = (d.kind == 1 && d.cin <= 256);        // digit GEMM with packed integer recombination
    // (the saturating 8-bit fast path of the fp32 epilogue survives only for residual blocks; plain 8-bit layers
    // whose parameters do not fit the fixed-point requantizer take the generic fp32 path)
#define QV2X_RUN(GG, DG, F8)                                                     \
    {                                                                            \
        RequantEpilogue<GG, DG, F8> e{};                                         \
        fill(e);                                                                 \
        return run_layer(ctx, e);                                                \
    }
    if (L->groups == 1) {
        if (sat8 && has_res) {
            RequantEpilogue<1, false, true, true> e{};
            fill(e);
            return run_layer(ctx, e);
        }
        QV2X_RUN(1, false, false)
    }
    if (digits) QV2X_RUN(3, true, false)
    if (sat8 && !has_res) QV2X_RUN(3, false, true)
    QV2X_RUN(3, false, false)
#undef QV2X_RUN
}

}  // extern "C"
