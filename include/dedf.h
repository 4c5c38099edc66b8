/*
 * dedf.h -- C ABI of libdedf.so: hand-written sm_100a CUDA kernels for the Diffusion-EDF
 * score-network hot path (MultiscaleScoreModel.forward / ScoreModelBase.sample).
 *
 * The reference (tomato1mule/diffusion_edf) is pure Python: it has no FFI / plugin boundary of
 * its own.  Its de-facto operator boundary is the set of third-party ops the hot path calls
 * (SURVEY.md section 8b); every entry point below names the reference call site (file:line under
 * /root/reference/diffusion_edf unless noted) whose arithmetic it replaces.  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host (plain int/float arrays that are
 *     copied into the launch parameters); all buffers are allocated and owned by the caller
 *   - returns 0 on success, a negative DEDF_ERR_* code otherwise; never throws, never allocates, never
 *     synchronises; work is enqueued on `stream`
 *   - features are fp32, row-major (N, F), e3nn "mul_ir" layout of an irreps triple
 *     (m0 x0e + m1 x1e + m2 x2e): [m0 scalars | m1 x 3 | m2 x 5]; `irr` arguments are int[3] = {m0,m1,m2}
 *   - graphs are CSR by destination: row_ptr[(n_seg * n_dst) + 1], edge_src[E], edge_dst[E] (int32);
 *     edges of destination d in segment s are [row_ptr[s*n_dst+d], row_ptr[s*n_dst+d+1]); sources ascend
 *   - the number of edges of a data-dependent graph lives on the device (`n_edges_dev`); `max_edges` only
 *     sizes the launch
 */
#ifndef DEDF_H
#define DEDF_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return codes of every entry point */
enum { DEDF_OK = 0, DEDF_ERR_ARG = -1 /* null pointer, bad size, misaligned operand */, DEDF_ERR_LAUNCH = -2 /* cudaGetLastError() after a launch */,
       DEDF_ERR_UNSUPPORTED = -3 /* irreps / widths outside the implemented family */ };

#define DEDF_MAX_SCALES 8
#define DEDF_MLP_MAX_LAYERS 4

#define DEDF_MLP_IN_ROWS 0   /* input rows given explicitly */
#define DEDF_MLP_IN_RBF 1    /* GaussianRadialBasisLayerFiniteCutoff(length)            (radial_func.py:231-278) */
#define DEDF_MLP_IN_FIELD 2  /* per-scale length encoder + edge_scalars_pre_linears     (graph_parser.py:180-183,
                                                                                         multiscale_tensor_field.py:225-234) */
#define DEDF_EPI_ACT 0       /* alpha logits + Gate                                     (graph_attention.py:233-246) */
#define DEDF_EPI_LIN 1       /* bias only                                               (graph_attention.py:237-239) */

/* ---- graph construction ------------------------------------------------------------------------------ */

/* torch_cluster.fps(src, ratio, random_start) for ONE batch segment (connectivity.py:62).
 * Selects m points greedily (arg-max of the running min squared distance, ties -> lowest index) starting from
 * `start` (or `*start_dev` when non-NULL: a device-side random start, so that the call stays capturable in a CUDA
 * graph); writes idx_base + index in selection order.  scratch_dist (n floats) is needed only for n > 16384. */
int dedf_fps(const float* x, int n, int m, int start, const long long* start_dev, int idx_base, long long* out_idx,
             float* scratch_dist, cudaStream_t stream);

/* torch_cluster.radius / radius_graph / the all-pairs meshgrid (graph_parser.py:339, :276-278;
 * connectivity.py:22, :42, :68-70), for n_scales source clouds at once (sources concatenated, cloud s =
 * [src_off_host[s], src_off_host[s+1]), radius r_host[s], r < 0 = all pairs).  Pass 1 writes counts
 * (n_scales * n_dst) and their exclusive scan row_ptr (+1 entry = E); pass 2 fills edge_src (flat source index)
 * and edge_dst.  excl_mode: 0 none; 1 drop src == excl[dst]; 2 drop src == dst; 3 drop excl[src] == dst.
 * max_nb caps the hits per (scale, dst) BEFORE the exclusion, like torch_cluster + the reference's filter.
 * capacity > 0 (edge buffers sized ahead of time, e.g. for CUDA-graph replay): row_ptr is clamped to `capacity`,
 * *n_edges_out = min(E, capacity) and *overflow |= 1 if E > capacity; capacity <= 0: exact CSR. */
int dedf_radius_count(const float* x_src, const float* x_dst, int n_dst, int n_scales, const int* src_off_host,
                      const float* r_host, const long long* b_src, const long long* b_dst, int excl_mode,
                      const long long* excl, int max_nb, int* counts, int* row_ptr, int capacity, int* n_edges_out,
                      int* overflow, cudaStream_t stream);
int dedf_radius_fill(const float* x_src, const float* x_dst, int n_dst, int n_scales, const int* src_off_host,
                     const float* r_host, const long long* b_src, const long long* b_dst, int excl_mode,
                     const long long* excl, int max_nb, const int* row_ptr, int* edge_src, int* edge_dst,
                     cudaStream_t stream);

/* Grid-hash variant of the radius search for ONE large source cloud (connectivity.py:22,42 at the fine scales; the C4
 * sweep): dedf_grid_build counting-sorts the sources into `n_buckets` (power of two, >= 32) hash buckets of a uniform grid
 * with cell edge 1.001 r (workspace: bucket_cnt[n_buckets], bucket_start[n_buckets+1], sorted_idx[n_src],
 * sorted_xyz[3 n_src]); the two passes then visit only the 27 surrounding cells of each destination.  Same arguments and
 * element-for-element the same CSR as dedf_radius_count / dedf_radius_fill with n_scales = 1. */
int dedf_grid_build(const float* x_src, int n_src, float r, int n_buckets, int* bucket_cnt, int* bucket_start,
                    int* sorted_idx, float* sorted_xyz, cudaStream_t stream);
int dedf_radius_grid_count(const float* x_src, int n_src, const float* x_dst, int n_dst, float r, int n_buckets,
                           const int* bucket_start, const int* sorted_idx, const float* sorted_xyz, const long long* b_src,
                           const long long* b_dst, int excl_mode, const long long* excl, int max_nb, int* counts,
                           int* row_ptr, int capacity, int* n_edges_out, int* overflow, cudaStream_t stream);
int dedf_radius_grid_fill(const float* x_src, int n_src, const float* x_dst, int n_dst, float r, int n_buckets,
                          const int* bucket_start, const int* sorted_idx, const float* sorted_xyz, const long long* b_src,
                          const long long* b_dst, int excl_mode, const long long* excl, int max_nb, const int* row_ptr,
                          int* edge_src, int* edge_dst, cudaStream_t stream);

/* ---- per-edge ---------------------------------------------------------------------------------------- */

/* Edge vector, length, o3.SphericalHarmonics(lmax=2, normalize=True, 'component'), non-scalar SH min-cut and
 * log soft cut-off (graph_parser.py:146-224; unet_feature_extractor.py:284-288).  logit may be NULL (UNet);
 * ns_hi <= 0 disables the min-cut; r_host[s] < 0 marks the infinite scale (logit 0). */
int dedf_edge_geom(const float* x_src, const float* x_dst, const int* edge_src, const int* edge_dst,
                   const int* n_edges_dev, int max_edges, int n_scales, const int* src_off_host, const float* r_host,
                   float ns_lo, float ns_hi, float* length, float* sh, float* logit, cudaStream_t stream);

typedef struct dedf_mlp_desc {
    int mode;                      /* DEDF_MLP_IN_* */
    const int* n_edges_dev;
    const float* x_in;             /* ROWS: (E, dims[0]) */
    const float* length;           /* RBF / FIELD: (E) */
    const float* rbf_mean; const float* rbf_std_logit; const float* rbf_weight_logit;   /* RBF: (dims[0]) each */
    float rbf_cutoff, rbf_offset;
    int n_scales, n_dst;           /* FIELD */
    const int* row_ptr; const int* edge_dst;
    const float* enc_mean[DEDF_MAX_SCALES]; const float* enc_std_logit[DEDF_MAX_SCALES];
    const float* enc_weight_logit[DEDF_MAX_SCALES];
    float enc_r[DEDF_MAX_SCALES];  /* >= 0: GaussianRadialBasis(max_val = r); < 0: sinusoidal(max_val = enc_max_r, n = enc_n) */
    float enc_max_r, enc_n;
    const float* enc_freq;         /* (dims[0]/2) sinusoidal frequencies exp(-k ln(n)/(half-1)), tabulated by the host in fp32 */
    const float* pre_w[DEDF_MAX_SCALES];   /* (dims[0], dims[1]): transposed length half of the pre-linear */
    const float* row_bias;         /* (n_scales, n_rb, dims[1]) from dedf_time_embed, or NULL (no context embedding: use b[0]) */
    int n_rb, rb_div;              /* row = min(edge_dst / rb_div, n_rb - 1) */
    int n_layers;
    int dims[DEDF_MLP_MAX_LAYERS + 1];
    const float* W[DEDF_MLP_MAX_LAYERS];      /* (dims[i], dims[i+1]) = torch Linear.weight^T */
    const float* b[DEDF_MLP_MAX_LAYERS];
    const float* ln_g[DEDF_MLP_MAX_LAYERS]; const float* ln_b[DEDF_MLP_MAX_LAYERS];
    int flags[DEDF_MLP_MAX_LAYERS];           /* bit0 LayerNorm, bit1 SiLU */
    const float* out_offset;       /* RadialProfile.offset (dims[n_layers]) or NULL */
    float* out;                    /* (E, dims[n_layers]) */
    /* tensor-core kernel only (dedf_edge_mlp_tc): weights pre-split into tf32 hi / lo parts and laid out as the MMA's B
     * operand: per layer, for each N block nb (N <= 256: one block; else N/2) and each 8-wide K chunk kc:
     * [hi: 2 x Nb x 4 floats | lo: same], element (n, kk) of a part at ((kk / 4) * Nb + n) * 4 + kk % 4
     * (diffusion_edf_b200/layers.py pack_tc).  pre_w_tc: the n_scales first-layer matrices of FIELD mode, back to back. */
    const float* W_tc[DEDF_MLP_MAX_LAYERS];
    const float* pre_w_tc;
    /* != 0: W_tc / pre_w_tc are the fp16 hi / lo packs (pack_tc(..., f16=True): 16-wide K chunks of 8-half groups) and the
     * products run on kind::f16 (three MMAs per K = 16; same accuracy as the tf32 split for |operands| < 65504, half the MMAs and
     * half the weight bytes streamed per tile).  Needs dims[i] % 16 == 0 for every layer input. */
    int tc_f16;
} dedf_mlp_desc;

/* RadialProfile MLP on the edge scalars (equiformer/radial_func.py:56-59) fused with the computation of its
 * input embedding. */
int dedf_edge_mlp(const dedf_mlp_desc* desc_host, int max_edges, cudaStream_t stream);

/* Same MLP on the tcgen05 tensor cores (kind::tf32 with the 3xTF32 hi/lo split: fp32-level accuracy), accumulator in TMEM,
 * weights streamed by TMA bulk copies, 128 edges per tile, one edge per epilogue thread.  Modes RBF and FIELD; in FIELD
 * mode the layers are [pre-linear (per scale, + time row bias, SiLU), RadialProfile...] in ONE launch (dims / b / ln / flags
 * describe all of them; layer 0 uses pre_w_tc).  Needs dims[i] % 8 == 0, dims[i] <= 128 for inputs and hidden layers,
 * output width <= 512 (split in two N blocks above 256). */
int dedf_edge_mlp_tc(const dedf_mlp_desc* desc_host, int max_edges, cudaStream_t stream);

/* gather message[src] (+ message_dst[dst]) -> DepthwiseTensorProduct 'uvu' with the 9 spherical harmonics
 * (equiformer/tensor_product_rescale.py:352-382 -> o3.TensorProduct) -> LinearRS block-diagonal linear ->
 * epilogue, the (E, 1568) tensor-product output never leaving shared memory (graph_attention.py:231-239).
 * mul1 = multiplicity of 1e (32: 64x0e+32x1e+16x2e, 16: 32x0e+16x1e+8x2e).
 * w: per-edge TP weights (E, numel) with w_stride = numel, or shared weights (numel) with w_stride = 0.
 * EPI_ACT: W0 = [sep_alpha | sep_act.lin 0e] (D0, MA + m0 + m1 + m2), bias0 likewise; writes logits (E,4)
 *          (SmoothLeakyReLU . alpha_dot + edge_logit) and out = Gate(...) (E, F).
 * EPI_LIN: W0 (D0, m0), bias0 (m0); writes out (E, F). */
int dedf_edge_tp_lin(int mul1, int epilogue, const float* x_src, const float* x_dst, int per_edge_x,
                     const int* edge_src, const int* edge_dst, const int* n_edges_dev, int max_edges,
                     const float* sh, const float* w, long long w_stride, const float* W0, const float* W1,
                     const float* W2, const float* bias0, const float* alpha_dot, const float* edge_logit,
                     float* logits, float* out, cudaStream_t stream);

/* dedf_edge_tp_lin with the EPI_ACT epilogue on the tcgen05 tensor cores (3xTF32, fp32 accumulators in TMEM): same
 * inputs, same outputs (graph_attention.py:231-246).  W_tc: the block-diagonal weights [sep_alpha | sep_act.lin] packed
 * per channel chunk as tf32 hi / lo B operands (diffusion_edf_b200/layers.py: pack_tp_act_tc documents the order);
 * bias0 (MA + m0 + m1 + m2) as for dedf_edge_tp_lin.  Gathers only (no per-edge x); x_src, x_dst, w 16-byte aligned,
 * w_stride % 4 == 0.  w_perm is a bit set: DEDF_TPACT_W_PERM = the columns of w are in the kernel's chunk-major order
 * (layers.tp_act_w_perm: the producer of w - the radial MLP - permutes its last layer), which turns 9 / 6 scalar gathers per lane
 * into 3 vector ones; DEDF_TPACT_F16 = W_tc is the fp16 hi / lo pack (layers.pack_tp_act_tc(..., f16=True)) and the products run
 * on kind::f16 (x = fp16 hi + fp16 lo, three MMAs per K = 16, fp32 accumulate: the accuracy of the tf32 split at half the
 * shared-memory traffic; operands must stay below 65504 in magnitude, larger ones come out as NaN). */
#define DEDF_TPACT_W_PERM 1
#define DEDF_TPACT_F16 2
int dedf_edge_tp_act_tc(int mul1, const float* x_src, const float* x_dst, const int* edge_src, const int* edge_dst,
                        const int* n_edges_dev, int max_edges, const float* sh, const float* w, long long w_stride,
                        int w_perm, const float* W_tc, const float* bias0, const float* alpha_dot, const float* edge_logit,
                        float* logits, float* out, cudaStream_t stream);

/* torch_scatter.scatter_logsumexp + exp + scatter(sum) (graph_attention.py:254-265): per destination and head
 * softmax of the logits over the incoming edges (all n_seg segments), out[d] = sum_e alpha[e, head(c)] val[e, c]. */
int dedf_segment_softmax_reduce(const int* row_ptr, int n_dst, int n_seg, const float* logits, const float* val,
                                int m0, int m1, int m2, float* out, cudaStream_t stream);

/* The value path of the attention, reassociated: out[d] = lin(sum_e alpha_{e,h} dtp(v_e, sh_e, w_shared)) +
 * bias * sum_e alpha_{e,h} with alpha = per-destination softmax of `logits` (x `post[e]` after the softmax when non-NULL).
 * Replaces sep_value (graph_attention.py:237-239: DepthwiseTensorProduct + LinearRS per EDGE) + scatter_logsumexp + exp +
 * scatter(sum) (:254-266): the linear layer runs once per destination.  v (E,F) and logits (E,4) must be 16-byte aligned. */
int dedf_value_reduce(int mul1, const int* row_ptr, int n_dst, int n_seg, const float* v, const float* sh, const float* logits,
                      const float* post, const float* wv, const float* V0, const float* V1, const float* V2, const float* vb,
                      float* out, cudaStream_t stream);

/* "K1": out[d] = sum_{e -> d} alpha[e, head(u)] * DTP(x[src_e], sh_e, w_e)   (N_dst, 1568 | 784), op-equivalent to
 * scatter(alpha * o3.TensorProduct(x[edge_src], sh, weight), edge_dst) of graph_attention.py:231-232,264-265.
 * sh_stride = 9: packed harmonics, plain loads.  sh_stride = 12: rows padded to 48 bytes, which lets every operand
 * (weight rows, gathered feature rows, harmonics, alphas) be staged by 1-D TMA bulk copies into a per-warp
 * shared-memory ring (the fast path). */
int dedf_edge_tp_reduce(int mul1, const float* x, const int* row_ptr, const int* edge_src, const float* sh,
                        int sh_stride, const float* w, const float* alpha, int n_dst, float* out, cudaStream_t stream);

/* ---- the front of one score-head evaluation, fused ------------------------------------------------------------
 * Pose transform of the query points (gnn_data.py:88-100, pcd_utils.py:55-81), multi-scale radius search against the
 * concatenated scene scales (graph_parser.py:336-345 torch_cluster.radius with max_num_neighbors, :272-286 all pairs for
 * r < 0; multiscale_tensor_field.py:236-247 concatenation order), CSR construction and the edge geometry
 * (graph_parser.py:146-224: length, l<=2 harmonics with the non-scalar min-cut, edge logits) in ONE launch.  Outputs are
 * bit-identical to dedf_query_transform (points) + dedf_radius_count / _fill + dedf_edge_geom.
 * Destinations are (pose t, query point q) -> row t * n_q + q; batch id of a destination = b_q[q].
 * capacity > 0: the edge buffers hold `capacity` edges; the CSR is clamped and *overflow raised when the true count exceeds it.
 * Workspace (caller-owned): counts (n_scales * n_dst), cta_sum (>= 148 ints), barrier (2 uints, ZEROED ONCE by the caller and
 * then left alone: a reusable device-wide barrier; one stream at a time may use a given barrier buffer).
 * Denoise loop (optional, replaces dedf_sample_advance): rows_all (n_scales, n_steps, rows_k) precomputed time rows and
 * *step (device step index) -> rows_cur (n_scales, 1, rows_k).  stage_early = 1 promises that x_src / b_src were written
 * before the PREVIOUS kernel of the stream started (static scene): they are then staged before the dependency wait. */
typedef struct dedf_head_front_desc {
    const float* Ts; int n_t;                  /* (n_t, 7) [qw qx qy qz x y z] fp32 */
    const float* qx; int n_q;                  /* (n_q, 3) query coordinates in the grasp frame */
    const float* x_src; const long long* b_src; const long long* b_q;   /* sources (sum N_s, 3); optional batch ids */
    int n_scales; int src_off[DEDF_MAX_SCALES + 1]; float r[DEDF_MAX_SCALES];
    int max_num_neighbors; float ns_lo, ns_hi; int capacity;
    float* x_dst;                              /* out (n_t * n_q, 3) */
    int* row_ptr; int* counts; int* edge_src; int* edge_dst;            /* out: row_ptr (n_scales * n_dst + 1) */
    float* length; float* sh; float* logit;    /* out (E), (E, 9), (E) */
    int* n_edges; int* overflow;               /* out: device scalar(s); overflow may be NULL */
    int* cta_sum; unsigned* barrier;           /* workspace */
    const int* step; int n_steps; const float* rows_all; float* rows_cur; int rows_k;
    int stage_early;
} dedf_head_front_desc;
int dedf_head_front(const dedf_head_front_desc* d, cudaStream_t stream);

/* ---- per-node ---------------------------------------------------------------------------------------- */

/* y = epilogue( LinearRS( [EquivariantLayerNormV2](x) ) ): equiformer/layer_norm.py:91-156,
 * equiformer/tensor_product_rescale.py:176-185 / :155-173 / :241-268, equiformer/fast_activation.py:210-224,
 * skip.py:13-35.  W_l (in.m_l, out.m_l) row-major; ln_w NULL = no norm; gate = 1 applies the SiLU/sigmoid Gate
 * (out.m0 = scalars + gates); y = (y + res) * res_scale when res != NULL. */
int dedf_node_linear(const float* x, int n, const int* irr_in_host, const int* irr_out_host, const float* W0,
                     const float* W1, const float* W2, const float* bias0, const float* ln_w, const float* ln_b,
                     float ln_eps, int gate, const float* res, float res_scale, float* y, cudaStream_t stream);

/* The per-node tail of an Equiformer block in ONE launch (replaces three dedf_node_linear calls):
 *     y1 = proj(x) + b_p (+ res1)                  graph_attention.py:118-121,268-272 + the block's first residual
 *     y  = y1 + fctp_2(Gate(fctp_1(LN(y1))))       gnn_block.py:51-57,207-216 / block.py:51-57,165-173
 * proj: irr_emb -> irr_emb; fctp_1: irr_emb -> irr_pre (pre-gate, m0 = scalars + gates); fctp_2: Gate(irr_pre) -> irr_emb.
 * Weight blocks (in.m_l, out.m_l) row-major as for dedf_node_linear; every multiplicity must be a positive multiple of 4 and
 * every pointer 16-byte aligned (else DEDF_ERR_UNSUPPORTED: the caller falls back to three dedf_node_linear launches).
 * Bit-identical to that un-fused sequence (same micro-kernel and summation order). */
typedef struct dedf_node_chain_desc {
    const float* x; int n;                     /* (n, dim(irr_emb)) attention output */
    int irr_emb[3], irr_pre[3];
    const float *P0, *P1, *P2, *pb;            /* proj weights / 0e bias (pb may be NULL) */
    const float* res1;                         /* optional (n, dim(irr_emb)) added to proj's output */
    const float *ln_w, *ln_b; float ln_eps;    /* EquivariantLayerNormV2 affine_weight (num_irreps) / affine_bias (m0) */
    const float *A0, *A1, *A2, *ab;            /* fctp_1 */
    const float *B0, *B1, *B2, *bb;            /* fctp_2 */
    float* y;                                  /* (n, dim(irr_emb)) */
} dedf_node_chain_desc;
int dedf_node_chain(const dedf_node_chain_desc* d, cudaStream_t stream);

/* Two independent LinearRS problems (no norm / gate / residual) with a common output irreps in ONE launch: the
 * linear_src / linear_dst pair at the head of every UNet block (block.py:149-153).  W_x = {W0, W1, W2} host arrays. */
int dedf_node_linear_pair(const float* x_a, int n_a, const int* irr_in_a_host, const float* const* W_a_host3, const float* bias_a, float* y_a,
                          const float* x_b, int n_b, const int* irr_in_b_host, const float* const* W_b_host3, const float* bias_b, float* y_b,
                          const int* irr_out_host, cudaStream_t stream);

/* KeypointExtractor.weight_post (keypoint_extractor.py:129-134, :186-190): y = act(Linear(SiLU(LayerNorm(x)))) * softplus(mult). */
int dedf_weight_post(const float* x, int n, int dim, const float* ln_g, const float* ln_b, const float* w, const float* b,
                     int use_sigmoid, const float* mult_logit, float* y, cudaStream_t stream);

int dedf_gather_rows(const float* x, const long long* idx, int n, int F, float* y, cudaStream_t stream);
int dedf_add_scale(const float* a, const float* b, float s, long long n, float* y, cudaStream_t stream);

/* ---- score head -------------------------------------------------------------------------------------- */

typedef struct dedf_time_desc {
    float max_time, enc_n;
    const float* enc_freq;         /* (enc_dim/2) frequencies exp(-k ln(n)/(half-1)), tabulated by the host in fp32 */
    int enc_dim, h_dim, e_dim, out_dim, n_scales;
    const float* W1[DEDF_MAX_SCALES]; const float* b1[DEDF_MAX_SCALES];   /* (enc_dim, h_dim) */
    const float* W2[DEDF_MAX_SCALES]; const float* b2[DEDF_MAX_SCALES];   /* (h_dim, e_dim) */
    const float* Wp[DEDF_MAX_SCALES]; const float* bp[DEDF_MAX_SCALES];   /* (e_dim, out_dim) */
} dedf_time_desc;

/* SinusoidalPositionEmbeddings + time MLPs (score_head.py:53-63,160-164) + the time half of
 * edge_scalars_pre_linears (multiscale_tensor_field.py:225-234): out (n_scales, n_t, out_dim). */
int dedf_time_embed(const dedf_time_desc* desc_host, const float* time, int n_t, float* out, cudaStream_t stream);

/* TransformPcd.forward (gnn_data.py:88-100): x' = q x q^-1 + t, f' = D(q) f  (wigner.py:257-283). */
int dedf_query_transform(const float* Ts, int n_t, const float* qx, const float* qf, int n_q, const int* irr_host,
                         float* x_out, float* f_out, cudaStream_t stream);

/* lin_vel_tp / ang_vel_tp (SeparableFCTP with shared weights), Gate, mean over the vectors, rotation by q^-1,
 * orbital term and weighted sum over the query points (score_head.py:192-209).  Arrays of 2 = (lin, ang).
 * Wd: the 9 path blocks of dtp.tp.weight in creation order, each transposed to [mul2][mul1]. */
int dedf_score_tp(const float* Ts, int n_t, const float* qf_rot, const float* key_f, const float* qx, const float* qw,
                  int n_q, const int* irr_host, const float* const* Wd_host2, const float* const* Wl0_host2,
                  const float* const* Wl1_host2, const float* const* bl_host2, int n_vec, float lin_mult,
                  float* ang_out, float* lin_out, cudaStream_t stream);

/* dedf_score_tp with the feature rotation D(q) psi of dedf_query_transform applied inside (qf = the UN-rotated query
 * features (n_q, F)) and, optionally, the Langevin step of dedf_pose_update fused behind it for a replayed denoise loop
 * (score_head.py:186-209 + score_model_base.py:174-199): when T64 != NULL, every pose's new state is integrated in float64
 * from row *counter of the device schedule `sched` (n_steps, 4) = [t, alpha_ang, alpha_lin, temperature], written to T64
 * (in place), to row *counter + 1 of `traj` and, cast to fp32, to T32 (which must be the `Ts` this call read);
 * the last CTA to finish advances *counter.  noise (n_steps, n_t, 6) or NULL = Philox as in dedf_pose_update.
 * ticket: one zero-initialised unsigned owned by the caller. */
typedef struct dedf_score_step_desc {
    const float* Ts; int n_t;
    const float* qf; const float* key_f; const float* qx; const float* qw; int n_q;
    int irr[3];
    const float* Wd[2]; const float* Wl0[2]; const float* Wl1[2]; const float* bl[2];
    int n_vec; float lin_mult;
    float* ang_out; float* lin_out;
    double* T64; const double* sched; int n_steps; int* counter; const double* noise; unsigned long long seed;
    const unsigned long long* seed_dev;        /* optional: the seed is read from device memory instead (graph replay) */
    double ang_mult, lin_mult_d; double* traj; float* T32; unsigned* ticket;
} dedf_score_step_desc;
int dedf_score_tp_step(const dedf_score_step_desc* d, cudaStream_t stream);

/* One annealed-Langevin step on SE(3) in float64 (score_model_base.py:178-193).  noise (n_t, 6) standard normals
 * or NULL: Philox4x32-10, seed = `seed`, subsequence = pose index, and `offset` = the STEP index -- every step owns a
 * disjoint window of 16 32-bit outputs of the pose's stream (a step consumes 12), so no two steps share a draw.
 * Optionally copies the new poses to traj_out (f64) and
 * T_f32_out (f32, the network input of the next step).
 * Graph-replay mode: dev_row (4 doubles [t, alpha_ang, alpha_lin, temperature], see dedf_sample_advance) overrides the
 * host scalars and dev_counter (device step index) selects the Philox step window, the noise rows (noise + step*n_t*6) and
 * the trajectory row (traj_out + (step+1)*n_t*7); the counter is incremented at the end of the call. */
int dedf_pose_update(double* T, int n_t, const float* ang, const float* lin, const double* noise,
                     unsigned long long seed, unsigned long long offset, double t, double ang_mult, double lin_mult,
                     double alpha_ang, double alpha_lin, double temperature, double* traj_out, float* T_f32_out,
                     const double* dev_row, int* dev_counter, cudaStream_t stream);

/* Loads row `*counter` of the device-resident schedule (n_steps, 4) = [t, alpha_ang, alpha_lin, temperature] into
 * cur_row and writes the fp32 time of the step to time_out[0] (the `time` input of dedf_time_embed), so that one denoise
 * step (score_model_base.py:146-199) is a parameter-free, replayable sequence of launches.  Optional: rows_all
 * (n_scales, n_steps, k) = the output of ONE dedf_time_embed launch over the whole schedule (known before the loop starts);
 * this step's rows are copied to rows_cur (n_scales, 1, k), which the edge MLP reads as its per-pose bias. */
int dedf_sample_advance(const double* sched, int n_steps, int* counter, float* time_out, double* cur_row,
                        const float* rows_all, float* rows_cur, int n_scales, int k, cudaStream_t stream);

/* library self-description: returns the compute capability the kernels were built for (100) */
/* EbmScoreModelHead.compute_energy tail (score_head_ebm.py:171-172): energy[t] = scale * sum_q w_q |key_f[t,q,:] - query_f[t,q,:]|^2 */
int dedf_ebm_energy(const float* key_f, const float* query_f, const float* qw, int n_t, int n_q, int F, float scale, float* out,
                    cudaStream_t stream);

/* ---- pre-processing in front of the path -------------------------------------------------------
 * voxel_filter (edf_interface/edf_interface/data/pcd_utils.py:123-152; preprocess.downsample, preprocess.py:69-80): voxel index =
 * trunc((p - min) / voxel_size), output sorted by the C-order ravelled index, features averaged, coordinates averaged
 * (center = 0) or voxel centres (center = 1).  dedf_bbox -> host reads mins / maxs (the reference syncs here too) and sizes the
 * dense grid (sx, sy, sz <= 2^27 cells); dedf_voxel_count fills key (n) / cnt (S, caller-zeroed) / off (S) / rank (S) and the
 * number of occupied voxels; dedf_voxel_reduce writes the (n_occupied, 3) / (n_occupied, F) outputs, summing every voxel's
 * points in ascending point-index order (deterministic, bit-identical to a sequential scatter). */
int dedf_bbox(const float* points, int n, float* mins, float* maxs, cudaStream_t stream);
int dedf_voxel_count(const float* points, int n, const float* mins, float voxel_size, int sx, int sy, int sz, int* key,
                     int* cnt_zeroed, int* off, int* rank, int* n_occupied, cudaStream_t stream);
int dedf_voxel_reduce(const float* points, const float* feats, int n, int F, const float* mins, float voxel_size, int sx, int sy,
                      int sz, const int* key, const int* cnt, const int* off, const int* rank, int* cursor_zeroed, int* sorted,
                      int center, float* out_points, float* out_feats, cudaStream_t stream);

/* ---- training path (un-fused primitives + their backward kernels) ---------------------------------------------
 * The reference trains through torch autograd over e3nn / torch_scatter ops (trainer.py:308-346 ->
 * score_model_base.py:41-107).  diffusion_edf_b200/autograd_ops.py wraps the pairs below in torch.autograd.Function so
 * that loss.backward() runs these kernels; torch itself only moves memory (cat / slice / reshape).  Parameter
 * gradients are ACCUMULATED (atomics) into caller-zeroed buffers. */

/* dW_l (in.m_l, out.m_l) and dbias (out.m0) of a block-diagonal linear y_l = W_l^T x_l (LinearRS, FCTP with 1x0e, nn.Linear =
 * irreps (K,0,0) -> (N,0,0)); dx is dedf_node_linear with the transposed weights.  tensor_product_rescale.py:176-185 */
int dedf_lin_wgrad(const float* x, const float* dy, int n, const int* irr_in_host, const int* irr_out_host, float* dW0,
                   float* dW1, float* dW2, float* dbias0, cudaStream_t stream);
/* EquivariantLayerNormV2 ('component', affine), equiformer/layer_norm.py:91-156; nn.LayerNorm is irreps (N,0,0). */
int dedf_ln_fwd(const float* x, int n, const int* irr_host, const float* w, const float* b, float eps, float* y, cudaStream_t stream);
int dedf_ln_bwd(const float* x, const float* g, int n, const int* irr_host, const float* w, float eps, float* dx, float* dw,
                float* db, cudaStream_t stream);
/* Gate (equiformer/fast_activation.py:210-224): irr_pre = pre-gate irreps (m0 = scalars + gates); plain SiLU / sigmoid: act_*. */
int dedf_gate_fwd(const float* pre, int n, const int* irr_pre_host, float* y, cudaStream_t stream);
int dedf_gate_bwd(const float* pre, const float* g, int n, const int* irr_pre_host, float* dpre, cudaStream_t stream);
int dedf_act_fwd(const float* x, long long n, int mode, float* y, cudaStream_t stream);      /* mode 0: SiLU, 1: sigmoid */
int dedf_act_bwd(const float* x, const float* g, long long n, int mode, float* dx, cudaStream_t stream);
/* DepthwiseTensorProduct 'uvu' (tensor_product_rescale.py:352-382): out (E, 49 mul1) in the sorted-irreps layout; w per edge
 * (w_stride = 15 mul1) or shared (w_stride = 0, dw accumulated). */
int dedf_dtp_fwd(int mul1, const float* x, const float* sh, const float* w, long long w_stride, int n_edges, float* out,
                 cudaStream_t stream);
int dedf_dtp_bwd(int mul1, const float* x, const float* sh, const float* w, long long w_stride, const float* g, int n_edges,
                 float* dx, float* dw, cudaStream_t stream);
/* The same tensor product for ANY even-parity l <= 2 irreps / harmonics degree, driven by a path table (creation order of
 * tensor_product_rescale.py:352-382): paths_host = n_paths rows of (l1, l2, lo, mul, w_off, ch_off); out (E, dim(irr_out)) in the sorted
 * simplified layout.  Forward only: the un-fused tensor field of irreps outside the fused kernels' family (BASELINE config C1). */
#define DEDF_DTP_MAX_PATHS 15
int dedf_dtp_generic_fwd(const float* x, const int* irr_in_host, const float* sh, const float* w, long long w_stride, int n_paths,
                         const int* paths_host, const int* irr_out_host, int n_edges, float* out, cudaStream_t stream);
int dedf_gather_rows_i32(const float* x, const int* idx, int n, int F, float* y, cudaStream_t stream);
int dedf_scatter_add_rows(const float* g, const void* idx, int idx_is_i64, int n, int F, float* out, cudaStream_t stream);
/* attention logits sum_k c SLReLU(pre[e,h,k]) alpha_dot[h,k] + edge_logit[e]   (graph_attention.py:241-246) */
int dedf_alpha_fwd(const float* pre, int n_edges, int ma, const float* alpha_dot, const float* edge_logit, float* logits,
                   cudaStream_t stream);
int dedf_alpha_bwd(const float* pre, int n_edges, int ma, const float* alpha_dot, const float* g, float* dpre, float* dalpha_dot,
                   cudaStream_t stream);
/* backward of dedf_segment_softmax_reduce (graph_attention.py:254-265) */
int dedf_softmax_reduce_bwd(const int* row_ptr, int n_dst, int n_seg, const float* logits, const float* val, const float* gout,
                            int m0, int m1, int m2, float* dlogits, float* dval, cudaStream_t stream);
/* Gaussian radial bases with learnable mean / std_logit / weight_logit (radial_func.py:168-278); mode 0: GaussianRadialBasis
 * (d = len * inv_span), mode 1: GaussianRadialBasisLayerFiniteCutoff (d = (len - offset) * inv_span, inner soft cut-off). */
int dedf_rbf_fwd(const float* len, int n_edges, int k, const float* mean, const float* std_logit, const float* weight_logit,
                 float offset, float inv_span, int mode, float* out, cudaStream_t stream);
int dedf_rbf_bwd(const float* len, int n_edges, int k, const float* mean, const float* std_logit, const float* weight_logit,
                 float offset, float inv_span, int mode, const float* g, float* dmean, float* dstd_logit, float* dweight_logit,
                 cudaStream_t stream);
/* SinusoidalPositionEmbeddings (radial_func.py:291-316): out[r] = [sin(x scale f_k) | cos(x scale f_k)] */
int dedf_sinusoid(const float* x, int n, int dim, const float* freq, float scale, float* out, cudaStream_t stream);
/* the 'uvu' score tensor product (score_head.py:123-139), un-fused: out (n, D0 + 3 D1); w = tp.weight in the reference layout */
int dedf_score_tp_fwd(const float* a, const float* b, const float* w, int n, const int* irr_host, float* out, cudaStream_t stream);
int dedf_score_tp_bwd(const float* a, const float* b, const float* w, int n, const int* irr_host, const float* g, float* da,
                      float* db, float* dw, cudaStream_t stream);
/* adjoint of dedf_query_transform w.r.t. the query features (accumulates over poses) */
int dedf_query_transform_bwd(const float* Ts, int n_t, int n_q, const int* irr_host, const float* g, float* dqf, cudaStream_t stream);
/* score_head.py:196-209 on the gated (n_t n_q, 1 + 3 n_vec) outputs of the lin / ang products */
int dedf_assemble_fwd(const float* Ts, int n_t, int n_q, int n_vec, const float* ylin, const float* yang, const float* qx,
                      const float* qw, float lin_mult, float* ang, float* lin, cudaStream_t stream);
int dedf_assemble_bwd(const float* Ts, int n_t, int n_q, int n_vec, const float* ylin, const float* yang, const float* qx,
                      const float* qw, float lin_mult, const float* gang, const float* glin, float* dylin, float* dyang,
                      float* dqw, cudaStream_t stream);

/* train-mode dropout: Philox mask (0 or 1/(1-p)); y = x * mask broadcast per attention head (mode 0, mask (n,4): nn.Dropout
 * on the attention weights, graph_attention.py:111-112), per irrep channel (mode 1, mask (n, m0+m1+m2):
 * EquivariantDropout, equiformer/drop.py:76-96) or per row (mode 2, mask (n,1)).  The backward of group_scale is group_scale on the gradient. */
int dedf_dropout_mask(unsigned long long seed, unsigned long long offset, long long n, float p, float* out, cudaStream_t stream);
int dedf_group_scale(const float* x, const float* mask, int n, const int* irr_host, int mode, float* y, cudaStream_t stream);
/* source-point attention (gnn_block.py:190-193 -> graph_attention.py:258-259: alpha_e *= w[src_e] after the softmax):
 * dedf_edge_gather_scalar builds the (E,1) factor (0 beyond *n_edges_dev), dedf_group_scale mode 2 applies it per row;
 * dedf_rowdot is the gradient w.r.t. the factor. */
int dedf_edge_gather_scalar(const float* w, const int* edge_src, const int* n_edges_dev, int max_edges, float* out, cudaStream_t stream);
int dedf_rowdot(const float* a, const float* b, int n, int F, float* out, cudaStream_t stream);

/* ---- position gradients: EbmScoreModelHead.forward (score_head_ebm.py:192-222) ----------------------------------------------
 * The reference's score of the energy-based head is torch.autograd.grad(-energy, T) through the tensor field; the adjoints of
 * the pieces that depend on the query coordinates (graph_parser.py:146-224 edge geometry, the length embeddings, the harmonics
 * operand of the depthwise tensor products) and the pull-back to the body frame are kernels of their own:
 *   dedf_dtp_bwd_sh     dsh (E, 9) of dedf_dtp_fwd
 *   dedf_rbf_bwd_len    dlen (E) of dedf_rbf_fwd;  dedf_sinusoid_bwd: dx (n) of dedf_sinusoid
 *   dedf_edge_geom_bwd  (g_len, g_sh, g_logit [may be NULL]) of dedf_edge_geom -> dx_dst (n_dst, 3), ACCUMULATED (zero it first)
 *   dedf_ebm_energy_bwd dkey / dquery (n_t n_q, F) of dedf_ebm_energy given g_energy (n_t)
 *   dedf_ebm_pose_grad  g_x (n_t n_q, 3), g_f (n_t n_q, F) = d logP / d(transformed coordinates, rotated features) ->
 *                       ang (n_t, 3) = ang_mult * right-trivialised rotational derivative, lin (n_t, 3) = lin_mult R^-1 d/dp */
int dedf_dtp_bwd_sh(int mul1, const float* x, const float* w, long long w_stride, const float* g, int n_edges, float* dsh,
                    cudaStream_t stream);
int dedf_rbf_bwd_len(const float* len, int n_edges, int k, const float* mean, const float* std_logit, const float* weight_logit,
                     float offset, float inv_span, int mode, const float* g, float* dlen, cudaStream_t stream);
int dedf_sinusoid_bwd(const float* x, int n, int dim, const float* freq, float scale, const float* g, float* dx, cudaStream_t stream);
int dedf_edge_geom_bwd(const float* x_src, const float* x_dst, const int* edge_src, const int* edge_dst, int n_edges, int n_scales,
                       const int* src_off_host, const float* r_host, float ns_lo, float ns_hi, const float* g_len, const float* g_sh,
                       const float* g_logit, float* dx_dst, cudaStream_t stream);
int dedf_ebm_energy_bwd(const float* key_f, const float* query_f, const float* qw, const float* g_energy, int n_t, int n_q, int F,
                        float scale, float* dkey, float* dquery, cudaStream_t stream);
int dedf_ebm_pose_grad(const float* Ts, int n_t, int n_q, const int* irr_host, const float* qx, const float* qf, const float* g_x,
                       const float* g_f, float ang_mult, float lin_mult, float* ang, float* lin, cudaStream_t stream);

/* ---- post-processing behind the path: collision-aware pre-place trajectory optimisation ---------------------
 * edf_interface/edf_interface/utils/collision_utils.py.  x (n_x, 3) scene cloud; y grasp cloud, (n_y, 3) shared by all poses
 * (y_pose_stride = 0) or (n_pose, n_y, 3) (y_pose_stride = 3 n_y); Ts (n_pose, 7) poses applied to y on the fly
 * (pcd_utils.transform_points, raw quaternion) or NULL when y is already in the scene frame.
 *   dedf_collision_check   _check_pcd_collision (:18-34): hit[pose] = 1 iff some point of the pose has a scene point with d^2 < r^2
 *   dedf_collision_energy  _pcd_energy (:40-110): energy[pose] = sum over the neighbours of cutoff_r / (|x - y|_1 + eps cutoff_r);
 *                          method 0 'knn' (the max_num_neighbors nearest, then |x - y|_1 <= cutoff_r), 1 'radius' (the first
 *                          max_num_neighbors in index order with d^2 < cutoff_r^2); grad (n_pose, 6) = d energy / d (rot xyz, trans xyz)
 *                          of an infinitesimal world-frame motion (the reference's autograd result), or NULL
 *   dedf_collision_step    _se3_adjoint_lie_grad + the update of _optimize_pcd_collision_once (:116-196):
 *                          Ts_out = Ts * exp(-dt cutoff_r diag(1,1,1,c,c,c) Ad^T grad)  (se3._exp_map / se3._multiply) */
int dedf_collision_check(const float* x, int n_x, const float* y, long long y_pose_stride, const float* Ts, int n_pose, int n_y, float r,
                         int* hit, cudaStream_t stream);
int dedf_collision_energy(const float* x, int n_x, const float* y, long long y_pose_stride, const float* Ts, int n_pose, int n_y,
                          float cutoff_r, int max_num_neighbors, float eps, int method, float* energy, float* grad, cudaStream_t stream);
int dedf_collision_step(const float* Ts, const float* grad, int n_pose, float dt, float cutoff_r, float* Ts_out, cudaStream_t stream);

int dedf_build_arch(void);

/* Self-test of the tcgen05 path (tc.cuh): D[128,N] = A[128,K] . B[N,K]^T on the tensor cores with the accumulator in TMEM;
 * n_split = 1: plain TF32 (10-bit mantissa), n_split = 3: 3xTF32 error-compensated split (fp32-level accuracy).
 * N % 16 == 0, 16 <= N <= 256, K % 8 == 0.  One CTA; used by tests/test_gpu_kernels.py. */
int dedf_tc_selftest(const float* A, const float* B, int N, int K, int n_split, float* D, cudaStream_t stream);

/* cudaAccessPolicyWindow on `stream`: keep [base, base + bytes) resident in L2 (persisting hits, streaming misses) for the
 * kernels launched afterwards; bytes = 0 removes the window.  No reference counterpart (used around dedf_edge_tp_reduce).
 * Unlike the kernel entry points this one is HOST-side configuration: it launches nothing, but it may raise the device's
 * persisting-L2 limit (cudaDeviceSetLimit) and must not be called during stream capture. */
int dedf_l2_persist(const void* base, long long bytes, cudaStream_t stream);

/* *flag |= 1 if the int64 arrays a and b (n words) differ anywhere.  Guard of a replayed CUDA graph: the plan of a forward
 * bakes in host values derived from the batch ids (FPS segments, query batch ids); a call with the same shapes but another
 * batch layout must re-plan instead of replaying stale segments. */
int dedf_flag_if_differs(const long long* a, const long long* b, long long n, int* flag, cudaStream_t stream);

/* Warm the L2 with the model's weights: issues prefetch.global.L2 over n device ranges (ptrs_dev[i], bytes_dev[i]).
 * The reference has no counterpart (its ~10^3 launches per forward re-read the weights through the cache hierarchy
 * implicitly); here the few-CTA kernels of the coarse UNet scales would otherwise pay DRAM latency per weight row. */
int dedf_prefetch_l2(const void* const* ptrs_dev, const long long* bytes_dev, int n, cudaStream_t stream);

/* Profiling aid: *slot = %globaltimer (ns) at this point of `stream` (capturable).  profiles/run_timeline.py enqueues it between the
 * kernels of a forward to read the in-graph timeline of the main and the geometry stream. */
int dedf_stamp(unsigned long long* slot, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEDF_H */
