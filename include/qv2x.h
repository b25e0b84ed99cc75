/*
 * qv2x.h -- C ABI of libqv2x.so: the B200 (sm_100a) fast path for QuantV2X's fully quantized
 * intermediate-fusion inference (pillars -> PointPillars front end -> quantized BEV backbone convs -> codebook
 * encode -> [indices on the wire / in peer memory] -> codebook decode -> ego-side max / attention fusion ->
 * detection heads -> post-processing).
 *
 * Every entry point replaces one PyTorch call site of the reference (paths relative to the reference
 * repository root); the reference has no FFI of its own -- the binding a maintainer adds is the
 * ctypes shim shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch types.  Pointers named d_* are DEVICE pointers; everything
 *     else is host memory that is only read during the call.
 *   - every function returns 0 on success or a negative qv2x_status; qv2x_last_error() returns a
 *     thread-local message.  Nothing throws, exits or synchronises the device unless stated.
 *   - forward calls are asynchronous on `stream` (a cudaStream_t passed as void*).  Handles are
 *     immutable after create, so forwards are re-entrant across streams given distinct buffers.
 *   - activations are uint8 NHWC ("pixel-major"): [n_img][H][W][channel stride]; activation zero
 *     points are 0 (every quantized activation on this path follows a ReLU, reference
 *     quant_layer.py:177-187 gives zp = 0 for a non-negative range).
 */
#ifndef QV2X_H_
#define QV2X_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QV2X_OK = 0,
    QV2X_ERR_INVALID = -1,     /* bad argument / unsupported shape */
    QV2X_ERR_CUDA = -2,        /* a CUDA runtime or driver call failed */
    QV2X_ERR_DEVICE = -3,      /* not an sm_100 device */
    QV2X_ERR_NOMEM = -4
} qv2x_status;

const char* qv2x_last_error(void);
int qv2x_version(void);
/* 0 when `device` is a compute-capability 10.x GPU (tcgen05/TMEM present). */
int qv2x_device_check(int device);
/* Number of kernels this library has launched since load (all threads); bench.py reports it. */
long long qv2x_launch_count(void);
/* Bring-up / profiling knobs for the igemm kernels (0 = normal operation): 1 skip the epilogue math and stores,
 * 2 skip MMA issue, 4 skip activation (A) loads, 8 skip weight (B) loads, 16 skip the output stores only (results are
 * garbage with any of these); 64 forces the float64 path of the codebook encoder (results unchanged). */
void qv2x_set_debug_flags(int flags);
/* Bring-up: when d_buf != NULL the conv kernels record clock64 stamps per CTA / tile / role into
 * d_buf[grid][32 tiles][16 slots] (int64, device memory owned by the caller); NULL switches it off. */
void qv2x_debug_trace(long long* d_buf);

/* ------------------------------------------------------------------------------------------------
 * One quantized layer = reference QuantModule.forward (opencood/quant/quant_layer.py:391-410) with
 * use_weight_quant = use_act_quant = True: fake-quant(weight) -> F.conv2d / F.conv_transpose2d (+bias)
 * -> folded-BN identity -> ReLU -> fake-quant(activation), restated on integers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct qv2x_layer qv2x_layer;

typedef struct {
    uint32_t struct_size;/* = sizeof(qv2x_layer_desc): a caller built against another revision of this header is
                            rejected instead of being read as garbage (same first field in every descriptor) */
    int kind;            /* 0: nn.Conv2d, 1: nn.ConvTranspose2d with kernel_size == stride, padding 0 */
    int cin, cout;
    int ksize;           /* conv: 1 or 3; transposed conv: == stride */
    int stride;          /* conv: 1 or 2; transposed conv: 1, 2 or 4 */
    int pad;             /* conv: zeros on every side (3x3 'same' = 1; ZeroPad2d(1)+padding 0 = 1) */
    int w_bits;          /* 2..8 (reference UniformAffineQuantizer.n_bits) */
    int relu;            /* activation_function is nn.ReLU */
    int n_in_groups;     /* 1, or 3 when the input is torch.cat of three tensors with different scales
                            (reference base_bev_backbone.py:111-112); cin is split evenly */
    float in_delta[3];   /* act_quantizer.delta of the tensor(s) feeding this layer */
    float out_delta;     /* this layer's act_quantizer.delta */
    float out_zero_point;/* this layer's act_quantizer.zero_point (0 after ReLU) */
    int out_bits;        /* act_quantizer.n_bits (<= 8) */
    int groups;          /* nn.Conv2d groups (0 or 1 = dense).  > 1: w_int is [cout][cin/groups][k][k] (the ResNeXt
                            convs of the pyramid backbone, resblock.py:94); conv only */
} qv2x_layer_desc;

/* w_int: the integer weight grid round(w/delta)+zp clamped to [0, 2^w_bits-1], in PyTorch layout
 *        ([cout][cin][k][k] for conv, [cin][cout][k][k] for transposed conv), one byte each.
 * w_delta / w_zero_point: per dim-0 channel (cout entries for conv, cin for transposed conv --
 *        reference quant_layer.py:325-335 quantizes along dim 0 for both).
 * bias: cout floats or NULL.  All host pointers. */
int qv2x_layer_create(const qv2x_layer_desc* desc, const uint8_t* w_int, const float* w_delta,
                      const float* w_zero_point, const float* bias, qv2x_layer** out);
void qv2x_layer_destroy(qv2x_layer* layer);
/* 1 if forward needs d_rowsum_in (uint8 x uint8 path with per-channel weight zero-points). */
int qv2x_layer_needs_rowsum(const qv2x_layer* layer);

/* d_x: [n_img][hi][wi][in_cstride] uint8, the layer's channels start at in_cbase.
 * d_rowsum_in: n_in_groups device pointers (host array), each [n_img][hi][wi] int32 = per-pixel sum of
 *        that group's input bytes; may be NULL when qv2x_layer_needs_rowsum() == 0.
 * d_y: [n_img][ho][wo][out_cstride] uint8, written at channel out_cbase.
 * d_rowsum_out: [n_img][ho][wo] int32, ACCUMULATED into (caller zeroes it), or NULL.
 * d_acc_dump: [n_groups][n_img*gemm_rows][n_cols] int32 zero-point-corrected accumulators, or NULL
 *        (test hook for the "int32 accumulators bit-exact" check). */
int qv2x_layer_forward(const qv2x_layer* layer, int n_img, int hi, int wi, const uint8_t* d_x, int in_cstride,
                       int in_cbase, const int32_t* const* d_rowsum_in, uint8_t* d_y, int out_cstride,
                       int out_cbase, int32_t* d_rowsum_out, int32_t* d_acc_dump, void* stream);
/* Residual-block forms (reference QuantBasicBlock / QuantBottleneck.forward, opencood/quant/quant_block.py:88-97,
 * 124-134; the convs built with disable_act_quant=True, :79-86, :113-122):
 *   - shortcut: `out += residual` before the block's ReLU and act_quantizer.  Either the block input on its
 *     quantizer's grid (d_res_u8 codes, value = res_delta * code) or the FP32 output of the downsample conv
 *     (d_res_f32); both NHWC over the layer's OUTPUT pixels, row pitch res_cstride elements, first channel res_cbase.
 *   - d_out_f32: the layer has no act_quantizer (downsample conv, single_head_i): y (after the ReLU if desc.relu) is
 *     written as FP32 NHWC with row pitch out_f32_cstride; d_y may be NULL and d_rowsum_out must be.
 * extra == NULL is qv2x_layer_forward. */
typedef struct {
    uint32_t struct_size;          /* = sizeof(qv2x_layer_extra) */
    const uint8_t* d_res_u8;
    const float* d_res_f32;
    float res_delta;
    int res_cstride, res_cbase;
    float* d_out_f32;
    int out_f32_cstride;
} qv2x_layer_extra;
int qv2x_layer_forward_ex(const qv2x_layer* layer, int n_img, int hi, int wi, const uint8_t* d_x, int in_cstride,
                          int in_cbase, const int32_t* const* d_rowsum_in, uint8_t* d_y, int out_cstride,
                          int out_cbase, int32_t* d_rowsum_out, int32_t* d_acc_dump, const qv2x_layer_extra* extra,
                          void* stream);
/* Output extent of a layer for a given input extent. */
int qv2x_layer_out_shape(const qv2x_layer* layer, int hi, int wi, int* ho, int* wo);

/* Per-pixel channel sums of a uint8 NHWC tensor: d_out[p] = sum_c d_x[p*cstride + cbase + c], c < c. */
int qv2x_rowsum_u8(const uint8_t* d_x, long long n_pixels, int cstride, int cbase, int c, int32_t* d_out,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Codebook compressor = reference UMGMQuantizer.encode / .decode
 * (opencood/models/sub_modules/codebook.py:330-343 with :106-131, :192-201, :231-239, :263-269).
 * The wire payload of an agent is levels*m byte planes of `rows` codes each.
 * ---------------------------------------------------------------------------------------------- */
typedef struct qv2x_codebook qv2x_codebook;

typedef struct {
    uint32_t struct_size;/* = sizeof(qv2x_codebook_desc) */
    int channel;         /* C: feature channels (256 on the V2X-Real path) */
    int m;               /* seg_num: codebooks per level, each over C/m channels */
    int levels;          /* len(dict_size) residual levels (3 in the reference models) */
    int k[4];            /* dict_size per level, multiple of 16, <= 256 (codes are bytes) */
} qv2x_codebook_desc;

/* codebooks[l]: [m][k[l]][C/m] floats (nn.Parameter _codebook of level l).
 * weights[l*6+h] / biases[l*6+h]: nn.Linear(C, C) weight [C][C] (out, in) and bias [C] of head h of level l, in
 * the reference's component order: 0 latentStageEncoder, 1 quantizationHead, 2 latentHead, 3 dequantizationHead,
 * 4 sideHead, 5 restoreHead.  latentHead and sideHead are NULL on the last level.  Host pointers. */
int qv2x_codebook_create(const qv2x_codebook_desc* desc, const float* const* codebooks, const float* const* weights,
                         const float* const* biases, qv2x_codebook** out);
void qv2x_codebook_destroy(qv2x_codebook* cb);

/* encode: d_feat [rows][feat_cstride] uint8 activation codes (first C channels used) with scale `delta`
 * (x = delta * q, zero-point 0) -> d_codes [levels][m][rows] uint8.  argmin ties resolve to the lowest index.
 * `delta` is the static scale of the shrinker's output quantizer: the per-column score scales delta * sc[j] are
 * cached in the handle for the last delta seen.  A call with a DIFFERENT delta rewrites that cache with a
 * synchronous copy -- the one exception to "handles are immutable after create": such a call must not overlap
 * encodes of the same handle still in flight on other streams (use one handle per scale instead). */
int qv2x_codebook_encode(const qv2x_codebook* cb, long long rows, const uint8_t* d_feat, int feat_cstride,
                         float delta, uint8_t* d_codes, void* stream);
/* decode: d_codes [levels][m][rows] -> d_out [rows][C] float32 (pixel-major). */
int qv2x_codebook_decode(const qv2x_codebook* cb, long long rows, const uint8_t* d_codes, float* d_out, void* stream);

/* decode only rectangles of the (agent-major) row grid: region a covers rows base_row[a] + y * pitch + x for
 * y in [rect[4a], rect[4a+1]), x in [rect[4a+2], rect[4a+3]).  d_codes planes are plane_stride rows apart; d_out is
 * the full-size [rows][C] buffer, only the listed rows are written.  Used when the ego stage is sharded over GPUs:
 * every GPU decodes just the source area its output tile samples from.  base_row / rect are HOST arrays. */
int qv2x_codebook_decode_regions(const qv2x_codebook* cb, long long plane_stride, int pitch, int n_regions,
                                 const long long* base_row, const int* rect, const uint8_t* d_codes, float* d_out,
                                 void* stream);

/* The descriptor the handle was created from (desc->struct_size must be set by the caller). */
int qv2x_codebook_desc_get(const qv2x_codebook* cb, qv2x_codebook_desc* desc);

/* Test hook: the folded tables the kernels use.  which = 0 digits (int8 [sum_l 3*m*k_l][C], level-major then
 * digit-major), 1 per-column scale (double), 2 per-column constant (double), 3 codeword cross terms (double),
 * 4 decode constant (float [C]), 5 decode tables (float, level-major then segment-major [k_l][C]). */
long long qv2x_codebook_folded_size(const qv2x_codebook* cb, int which);
int qv2x_codebook_folded_copy(const qv2x_codebook* cb, int which, void* host_buf);

/* ------------------------------------------------------------------------------------------------
 * Ego-side fusion = reference MaxFusion / AttFusion .forward for one frame
 * (opencood/models/fuse_modules/fusion_in_one.py:87-151), including the bilinear warp of every agent
 * (the ego too) into the ego frame (warp_affine_simple, torch_transformation_utils.py:323-332).
 * d_feat: [n_agents][H][W][C] float32 pixel-major, agent 0 = ego.
 * d_affine: DEVICE [n_agents][2][3] float32 = normalize_pairwise_tfm(...)[b][0, :n] (ego row of the pairwise
 *         matrices, opencood/utils/transformation_utils.py:68-92); on the device so that a captured CUDA graph
 *         picks up new poses every frame.  mode 0 = max, 1 = attention (ego query row).
 * d_out:  [H][W][C] float32 pixel-major.
 * ---------------------------------------------------------------------------------------------- */
int qv2x_fuse(int mode, int n_agents, int H, int W, int C, const float* d_feat, const float* d_affine, float* d_out,
              void* stream);
/* Same, restricted to the output tile rows [y0, y1) x columns [x0, x1); d_out is compact [(y1-y0)*(x1-x0)][C]. */
int qv2x_fuse_tile(int mode, int n_agents, int H, int W, int C, const float* d_feat, const float* d_affine,
                   float* d_out, int y0, int y1, int x0, int x1, void* stream);
/* uint8 codes -> float32 values fl(delta * code), same (NHWC) layout; n elements, a multiple of 16. */
int qv2x_dequant_u8(const uint8_t* d_x, long long n, float delta, float* d_out, void* stream);
/* Score-weighted fusion of one pyramid level = reference weighted_fuse (opencood/models/fuse_modules/
 * pyramid_fuse.py:17-62) with the score preparation of QuantPyramidFusion.forward_collab
 * (opencood/quant/quant_block.py:516-539) folded in.  d_score: [n_agents][H][W] float32; score_is_logit = 1: the
 * occupancy logits of single_head_i, the kernel applies sigmoid(.) + 1e-4; 0: ready-made scores (e.g. after the
 * camera crop mask).  Scores are warped like the features; a warped score of exactly 0 excludes the agent;
 * out = sum_j softmax_j(score_j) x_j, or 0 where every agent is excluded.  Layouts as qv2x_fuse. */
int qv2x_fuse_weighted(int n_agents, int H, int W, int C, const float* d_feat, const float* d_score,
                       int score_is_logit, const float* d_affine, float* d_out, void* stream);
/* The same on the level's uint8 codes (scale delta, zero-point 0; C a multiple of 4): equals qv2x_dequant_u8 followed by
 * qv2x_fuse_weighted bit for bit, without the FP32 copy of every agent's map. */
int qv2x_fuse_weighted_u8(int n_agents, int H, int W, int C, const uint8_t* d_feat_u8, float delta,
                          const float* d_score, int score_is_logit, const float* d_affine, float* d_out,
                          void* stream);

/* Detection heads = the three 1x1 convs cls_head / reg_head / dir_head (reference
 * heter_model_baseline_mc.py:137-142) concatenated along the output channel; weights are the de-quantized
 * fake-quant weights (W-quant only, FP32 activations: quant_model.py:129-136).
 * w: HOST [cout][cin], bias HOST [cout] or NULL.  d_x [pixels][cin] pixel-major -> d_out [cout][pixels]
 * (= preds_tensor in NCHW for one frame).  cin <= 256, a multiple of 4; cout is arbitrary (72-column output chunks of
 * one launch: the FP32 1x1 / transposed convs of the pyramid path use the same GEMM). */
typedef struct qv2x_heads qv2x_heads;
int qv2x_heads_create(int cin, int cout, const float* w, const float* bias, qv2x_heads** out);
void qv2x_heads_destroy(qv2x_heads* heads);
int qv2x_heads_forward(const qv2x_heads* heads, long long pixels, const float* d_x, float* d_out, void* stream);
/* Same, for a tile of a larger map: the input is compact (pixels = tile_h * tile_w, row-major), output o of pixel
 * (ty, tx) is stored at d_out[o * out_pixels + ty * out_w + tx].  d_out may point INTO another GPU's peer-mapped
 * buffer (multi-GPU: every rank writes its tile of the head maps straight into the ego rank's result). */
int qv2x_heads_forward_tile(const qv2x_heads* heads, long long pixels, const float* d_x, float* d_out, int tile_w,
                            long long out_w, long long out_pixels, void* stream);
/* The same GEMM as a transposed conv with kernel = stride on an FP32 map, with the ReLU + activation quantizer and the
 * pixel shuffle in its epilogue (the deblocks of the pyramid path on the fused FP32 level maps: QuantModule(
 * ConvTranspose2d(cin, c, s, stride=s)) + ReLU + act quantizer, opencood/quant/quant_block.py:441-458).  `heads` was
 * created with cout = s*s*c rows ordered (dy, dx, channel); d_x [in_h*in_w][cin] float32; output pixel
 * (y*s + dy, x*s + dx), channel out_cbase + ch of the uint8 NHWC buffer d_out_u8 [in_h*s][in_w*s][out_cstride] gets
 * clamp(rint(value / out_delta), 0, 255) -- the codes of qv2x_heads_forward + qv2x_quantize_nchw_to_nhwc_u8, bit for
 * bit.  c, out_cbase and out_cstride must be multiples of 8. */
int qv2x_heads_forward_deconv_u8(const qv2x_heads* heads, int in_h, int in_w, const float* d_x, int stride,
                                 float out_delta, uint8_t* d_out_u8, int out_cstride, int out_cbase, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ego stage of an attention-fusion frame in one kernel: code planes in, head maps out.  Replaces the chain
 * UMGMQuantizer.decode (opencood/models/sub_modules/codebook.py:192-201, 263-269) -> warp_affine_simple
 * (opencood/models/sub_modules/torch_transformation_utils.py:323-332) -> AttFusion.forward
 * (opencood/models/fuse_modules/fusion_in_one.py:126-151) -> cls/reg/dir heads
 * (opencood/models/heter_model_baseline_mc.py:137-142), i.e. qv2x_codebook_decode + qv2x_fuse(mode 1) +
 * qv2x_heads_forward, without materialising any feature map: decode, warp and the heads are linear in the codeword
 * tables, so the attention scores come from the Gram matrix of the decode tables and the head maps from the tables
 * pushed through the head weights (both folded at create time in float64).  Floating-point tier: agrees with the
 * three-kernel chain to fp32 rounding (different summation order).
 * w: HOST [cout][C] head weights (as qv2x_heads_create), bias HOST [cout] or NULL.
 * Supported when levels*m is 1, 2, 3, 4 or 6, cout <= 72 and the [sum_l m*k_l][72] head table fits in shared memory
 * (qv2x_ego_att_supported returns 1); other configurations and max fusion use the three-kernel chain.
 * forward: d_codes [levels*m][plane_stride] uint8, agent a's rows at a*H*W (agent 0 = ego); d_affine DEVICE
 * [n_agents][2][3] as qv2x_fuse; d_out [cout][H*W] float32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct qv2x_ego_att qv2x_ego_att;
int qv2x_ego_att_supported(const qv2x_codebook* cb, int cout);
int qv2x_ego_att_create(const qv2x_codebook* cb, int cout, const float* w, const float* bias, qv2x_ego_att** out);
void qv2x_ego_att_destroy(qv2x_ego_att* h);
int qv2x_ego_att_forward(const qv2x_ego_att* h, int n_agents, int H, int W, const uint8_t* d_codes,
                         long long plane_stride, const float* d_affine, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decode + first 1x1 conv + activation quantizer of the pyramid model's ego stage, folded over the codeword tables
 * (SURVEY 8(f)-2).  Replaces UMGMQuantizer.decode (opencood/models/sub_modules/codebook.py:192-201, 263-269)
 * followed by conv1 (+ folded BN, ReLU, act quantizer) of the first ResNeXt bottleneck of PyramidFusion
 * (opencood/models/sub_modules/resblock.py:67-122 under QuantBottleneck, opencood/quant/quant_block.py:100-134):
 *   out[row][c] = clamp(rint((bias[c] + sum_k w[c][k] * decode(codes)[row][k]) / out_delta), 0, 255)
 * w: HOST [cout][C] de-quantized fake-quant weights, bias HOST [cout] or NULL; cout a multiple of 4.
 * forward: d_codes [levels*m][plane_stride] uint8 -> d_out [rows][cout] uint8 (pixel-major, zero-point 0) and, when
 * d_rowsum != NULL, the per-row sums of the output codes (the next layer's zero-point term).
 * Supported when levels*m <= 4 and the [sum_l m*k_l][cout] folded table fits in shared memory.
 * ---------------------------------------------------------------------------------------------- */
typedef struct qv2x_decode_linear qv2x_decode_linear;
int qv2x_decode_linear_supported(const qv2x_codebook* cb, int cout);
int qv2x_decode_linear_create(const qv2x_codebook* cb, int cout, const float* w, const float* bias, float out_delta,
                              qv2x_decode_linear** out);
void qv2x_decode_linear_destroy(qv2x_decode_linear* h);
int qv2x_decode_linear_forward(const qv2x_decode_linear* h, long long rows, const uint8_t* d_codes,
                               long long plane_stride, uint8_t* d_out, int32_t* d_rowsum, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PointPillars front end: pillars -> decorated points -> quantized Linear(10 -> 64) ->
 * pre-ReLU activation quantizer -> ReLU -> block activation quantizer -> max over the pillar's points -> scatter
 * into the uint8 NHWC BEV map, in one kernel.  Replaces QuantPointPillar.forward
 * (opencood/quant/quant_block.py:611-630, 666-741) + PointPillarScatter.forward
 * (opencood/models/sub_modules/point_pillar_scatter.py:19-75).
 * ---------------------------------------------------------------------------------------------- */
typedef struct qv2x_pillar qv2x_pillar;
typedef struct {
    uint32_t struct_size;  /* = sizeof(qv2x_pillar_desc) */
    int n_feat;            /* decorated point features: 10 (use_absolute_xyz, no distance) */
    int cout;              /* PFN output channels: 64 */
    int max_points;        /* points per pillar: 32 */
    int nx, ny;            /* BEV grid (cells along x / y) */
    float voxel_size[3];   /* voxel_x, voxel_y, voxel_z */
    float offset[3];       /* x_offset, y_offset, z_offset (cell centre of cell 0) */
    int has_pre_quant;     /* the Linear's own activation quantizer (applied before the ReLU) is active */
    float pre_delta, pre_zero_point;
    int pre_bits;
    float out_delta, out_zero_point;   /* the PFN block's post-ReLU quantizer = the BEV grid's scale */
    int out_bits;
} qv2x_pillar_desc;
/* w_hat: HOST float [cout][n_feat], the fake-quantized (de-quantized) weights with BN folded, exactly the tensor the
 * reference multiplies with (weight_quantizer(weight)); bias: HOST float [cout] or NULL. */
int qv2x_pillar_create(const qv2x_pillar_desc* desc, const float* w_hat, const float* bias, qv2x_pillar** out);
void qv2x_pillar_destroy(qv2x_pillar* p);
/* d_points float [n_pillars][32][4] (padded points zero), d_coords int32 [n_pillars][4] = (batch, z, y, x),
 * d_num_points int32 [n_pillars]; d_bev uint8 [batch][ny][nx][cout] is cleared and filled (empty cells = code 0). */
int qv2x_pillar_forward(const qv2x_pillar* p, int n_pillars, const float* d_points, const int* d_coords,
                        const int* d_num_points, int batch, uint8_t* d_bev, void* stream);
/* The same, also writing d_rowsum int32 [batch][ny][nx] = per-cell sum of the 64 codes (cleared, then filled): the
 * input row sums qv2x_plan_forward_rs takes, so that the plan need not scan the (mostly empty) map again. */
int qv2x_pillar_forward_rs(const qv2x_pillar* p, int n_pillars, const float* d_points, const int* d_coords,
                           const int* d_num_points, int batch, uint8_t* d_bev, int32_t* d_rowsum, void* stream);
/* Serving loops keep the map all-zero BETWEEN frames instead of clearing 9 MB per agent (95 % of it already zero) before
 * every frame: qv2x_pillar_scatter is qv2x_pillar_forward_rs without the clear (the caller guarantees d_bev and
 * d_rowsum are zero), qv2x_pillar_clear zeroes exactly the cells (and row sums) of the given pillars once the map has
 * been consumed.  scatter -> consumer -> clear leaves the buffers zero again. */
int qv2x_pillar_scatter(const qv2x_pillar* p, int n_pillars, const float* d_points, const int* d_coords,
                        const int* d_num_points, int batch, uint8_t* d_bev, int32_t* d_rowsum, void* stream);
int qv2x_pillar_clear(const qv2x_pillar* p, int n_pillars, const int* d_coords, int batch, uint8_t* d_bev,
                      int32_t* d_rowsum, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Detection post-processing on the GPU: sigmoid / max over classes / score threshold, box decode
 * against the anchors, BEV corners, rotated NMS (top-k by score, greedy), range mask.  Replaces
 * VoxelPostprocessor3Heads.post_process (opencood/data_utils/post_processor/voxel_postprocessor_3heads.py:318-477)
 * with nms_rotated (opencood/utils/box_utils_mc.py:665-710), which the reference runs on the CPU with shapely.
 * Anchors are generated on the fly: class c, rotation r, cell (y, x) has centre
 * (anchor_x0[c] + x * anchor_dx[c], anchor_y0[c] + y * anchor_dy[c], anchor_z[c]), size anchor_hwl[c], yaw
 * anchor_rot[r]; anchor index = ((y * W + x) * n_classes + c) * n_rotations + r, as the reference flattens them.
 * ---------------------------------------------------------------------------------------------- */
typedef struct qv2x_postprocess qv2x_postprocess;
typedef struct {
    uint32_t struct_size;           /* = sizeof(qv2x_postprocess_desc) */
    int H, W;                       /* head-map size */
    int n_classes, n_rotations;     /* anchors per cell = n_classes * n_rotations */
    double anchor_x0[4], anchor_y0[4], anchor_dx[4], anchor_dy[4], anchor_z[4];
    double anchor_hwl[4][3];
    double anchor_rot[4];
    double score_threshold;
    float nms_threshold;
    double range_lo[2], range_hi[2];   /* a box is kept only if all four BEV corners lie inside */
    int max_candidates;             /* capacity for anchors above the score threshold (excess is dropped) */
    int top;                        /* candidates that enter NMS (reference: 1000; at most 1024) */
} qv2x_postprocess_desc;
int qv2x_postprocess_create(const qv2x_postprocess_desc* desc, qv2x_postprocess** out);
void qv2x_postprocess_destroy(qv2x_postprocess* p);
/* d_preds float32 [n_cls*A + 7*A + ...][H*W] channel-major head maps (cls first, then reg), A = anchors per cell.
 * Outputs (device, capacity `top` boxes): d_corners double [K][4][2], d_scores double [K], d_labels int [K]
 * (1-based class), d_boxes double [K][7] (x, y, z, h, w, l, yaw), *d_n_out = K, in NMS pick order.
 * d_n_candidates (nullable): anchors that passed the score threshold (may exceed max_candidates). */
int qv2x_postprocess_forward(const qv2x_postprocess* p, const float* d_preds, double* d_corners, double* d_scores,
                             int* d_labels, double* d_boxes, int* d_n_out, int* d_n_candidates, void* stream);

/* Multi-GPU exchange of the code planes without a collective: stores d_local ([planes][rows_local] bytes) into the
 * code buffer of every peer: peer p receives plane i at peer_bases[p] + i * dst_plane_stride + dst_row0.
 * peer_bases is a HOST array of n_peers (<= 8) device pointers (peer-mapped; this rank's own buffer included).
 * Replaces the all-gather of reference-side int64 code lists (heter_pyramid_collab_codebook_mc_encdec.py:120-123
 * keeps them in host lists; the reference has no multi-GPU path). */
int qv2x_push_planes(const uint8_t* d_local, int planes, long long rows_local, long long dst_plane_stride,
                     long long dst_row0, void* const* peer_bases, int n_peers, void* stream);

/* The all-to-all form, for frame-batched serving (rank r holds the codes of ITS agents for n_peers consecutive frames,
 * frame f is fused on rank f): d_local is [planes][n_peers * rows_per_peer] bytes with frame-major rows; peer p
 * receives rows [p * rows_per_peer, (p + 1) * rows_per_peer) of plane i at
 * peer_bases[p] + i * dst_plane_stride + dst_row0. */
int qv2x_scatter_planes(const uint8_t* d_local, int planes, long long rows_per_peer, long long dst_plane_stride,
                        long long dst_row0, void* const* peer_bases, int n_peers, void* stream);

/* Layout / quantization converters for the module boundaries of the drop-in wrappers (the reference
 * passes float32 NCHW between modules; the kernels work on uint8 / float32 pixel-major tensors).
 * quantize: q = clamp(rint(x / delta) + zp, 0, 2^bits-1) (reference quant_layer.py:132-133). */
int qv2x_quantize_nchw_to_nhwc_u8(const float* d_x, int n, int c, long long pixels, float delta, float zero_point,
                                  int bits, uint8_t* d_y, int out_cstride, int out_cbase, void* stream);
int qv2x_dequant_nhwc_u8_to_nchw_f32(const uint8_t* d_x, int n, int c, long long pixels, float delta,
                                     float zero_point, float* d_y, void* stream);
int qv2x_nchw_to_nhwc_f32(const float* d_x, int n, int c, long long pixels, float* d_y, void* stream);
int qv2x_nhwc_to_nchw_f32(const float* d_x, int n, int c, long long pixels, float* d_y, void* stream);

/* ------------------------------------------------------------------------------------------------
 * A plan = the quantized BaseBEVBackbone + DownsampleConv of one modality as a fixed launch sequence
 * (reference QuantBaseBEVBackbone.forward quant_block.py:280-303 + QuantDownsampleConv.forward :583-586).
 * Buffers are numbered: 0 is the caller's input, n_bufs-1 the caller's output, the rest live in the
 * caller-provided workspace.  Every step runs one layer from (in_buf, in_cbase) to (out_buf, out_cbase);
 * several steps may write disjoint channel slices of one buffer (that is how torch.cat disappears).
 * ---------------------------------------------------------------------------------------------- */
typedef struct qv2x_plan qv2x_plan;
typedef struct {
    const qv2x_layer* layer;
    int in_buf, in_cbase;
    int out_buf, out_cbase;
} qv2x_plan_step;

/* buf_channels[b]: channel stride (bytes per pixel) of buffer b.  Layers must outlive the plan. */
int qv2x_plan_create(const qv2x_plan_step* steps, int n_steps, const int* buf_channels, int n_bufs, qv2x_plan** out);
void qv2x_plan_destroy(qv2x_plan* plan);
int qv2x_plan_out_shape(const qv2x_plan* plan, int H, int W, int* ho, int* wo, int* channels);
int qv2x_plan_workspace_bytes(const qv2x_plan* plan, int n_img, int H, int W, size_t* bytes);
/* d_in [n_img][H][W][buf_channels[0]] uint8 -> d_out [n_img][ho][wo][buf_channels[last]] uint8.
 * dump_step / d_acc_dump: optional accumulator dump of one step (see qv2x_layer_forward), -1 / NULL to disable. */
int qv2x_plan_forward(const qv2x_plan* plan, int n_img, int H, int W, const uint8_t* d_in, uint8_t* d_out,
                      void* d_workspace, size_t workspace_bytes, int dump_step, int32_t* d_acc_dump, void* stream);
/* The same with the per-pixel channel sums of the input supplied by its producer (d_in_rowsum int32 [n_img][H][W] over
 * ALL buf_channels[0] channels, e.g. from qv2x_pillar_forward_rs); NULL = compute them here (qv2x_plan_forward). */
int qv2x_plan_forward_rs(const qv2x_plan* plan, int n_img, int H, int W, const uint8_t* d_in, const int32_t* d_in_rowsum,
                         uint8_t* d_out, void* d_workspace, size_t workspace_bytes, int dump_step, int32_t* d_acc_dump,
                         void* stream);
/* Copy of the descriptor a layer was created with. */
int qv2x_layer_desc_get(const qv2x_layer* layer, qv2x_layer_desc* out);

/* ------------------------------------------------------------------------------------------------
 * Measurement aid: the raw tcgen05.mma kind::i8 rate of the device (dense int8 TOP/s, operands resident in shared
 * memory, one CTA per SM) and the SM clock seen while it ran.  The denominator of the conv kernels' roofline.
 * Synchronises `stream`.
 * ---------------------------------------------------------------------------------------------- */
int qv2x_int8_mma_peak(double* tops, double* sm_mhz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QV2X_H_ */
