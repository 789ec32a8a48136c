/* vbfem.h -- C ABI of libvbfem.so: B200 (sm_100a) batched Cook's-membrane FEM
 * forward + adjoint, the hot path inside the variational-Bayes ELBO.
 *
 * The upstream project (nfeng2022/Variational-Bayesian-Inference-for-
 * Computational-Mechanics) is pure Python and has no FFI; the seam this
 * library sits behind is the Python callable
 *     MeasurementData.fem_fh_fun_loop_rev(x[N,2]) -> (y[N,2], h[N,2])
 *     (src/data_generation_2sam_more_loss.py:169-192)
 * and, one level down, FemSolver.fea_solution (src/fem_solver_tf.py:13-73).
 * Every entry point cites the reference code it replaces.  All paths below
 * are relative to the upstream repository.
 *
 * Conventions
 *   - plain C types only; 0 = success, negative = error (vbfem_last_error()).
 *   - *_dev pointers are device pointers on the handle's GPU (obtained from
 *     torch / DLPack by the Python wrapper); *_host pointers are host memory.
 *   - everything is float64, row-major; index arrays handed in by the caller
 *     use the reference's 1-based numbering.
 *   - calls are stream-ordered on `stream` (a cudaStream_t passed as void*,
 *     NULL = default stream) and do not synchronise, except the *_host
 *     variants, vbfem_status and vbfem_create/destroy.
 *   - one handle per GPU; calls on one handle are not re-entrant.
 */
#ifndef VBFEM_H
#define VBFEM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vbfem_handle vbfem_t;

/* Sample-independent model description = the slice of
 * fem_preprocess.PreProcessing.model_data the hot path reads
 * (src/fem_preprocess.py:114-443, model_property_cards.py:25-29) plus the
 * MeasurementData class attributes (src/data_generation_2sam_more_loss.py:16-21,
 * main_custom_training.py:32-38). */
typedef struct vbfem_mesh {
    int32_t nnodes;          /* mesh_info['nnodes'] */
    int32_t nele;            /* mesh_info['nele'] */
    const double *coord;     /* [nnodes][2] = mesh_info['coord'][:, 1:3] */
    const int32_t *ien;      /* [nele][4]  = dof_info['IEN'], 1-based node ids */
    int32_t nfree;           /* dof_info['nfree'] */
    const int32_t *free_dof; /* [nfree] = dof_info['free_dof'], 1-based, dof = 2(n-1)+c+1 */
    const double *pf;        /* [nfree] = loading['Pf'] (dense) */
    double thk;              /* section[0]['thk'] */
    int32_t obs_node;        /* MeasurementData.node_id (1-based) */
    int32_t obs_ele;         /* MeasurementData.ele_id (1-based) */
    int32_t obs_gp[2];       /* MeasurementData.nipt_id (1-based Gauss points) */
    double theta_mean[2];    /* MeasurementData.theta_mean */
    double theta_std[2];     /* MeasurementData.theta_std */
} vbfem_mesh;

/* Index of the integers returned by vbfem_info(). */
enum {
    VBFEM_INFO_NFREE = 0,     /* n: order of the banded system */
    VBFEM_INFO_HALF_BW = 1,   /* b: half bandwidth under the internal numbering */
    VBFEM_INFO_NDOF = 2,
    VBFEM_INFO_NELE = 3,
    VBFEM_INFO_NCOLORS = 4,   /* element colours used by the atomics-free assembly */
    VBFEM_INFO_BAND_IN_SMEM = 5,
    VBFEM_INFO_SMEM_BYTES = 6,
    VBFEM_INFO_CTAS_PER_SM = 7,
    VBFEM_INFO_NUM_SMS = 8,
    VBFEM_INFO_BLOCK_THREADS = 9,
    VBFEM_INFO_KERNEL_VARIANT = 10, /* 0 = generic per-column kernel, 2 = on-chip two-front kernel,
                                       3 = blocked panel kernel (wide bands, factor streamed to HBM),
                                       4 = warp-per-sample kernel (narrow bands, window in registers) */
    VBFEM_INFO_TWIST_ROW = 11,      /* first middle row of the twisted factorisation */
    VBFEM_INFO_PANEL_BLOCKS = 12,   /* panel / warp kernel: block half bandwidth (8x8 blocks below the diagonal block) */
    VBFEM_INFO_PANEL_RING = 13,     /* panel / warp kernel: capacity of the element-matrix ring */
    VBFEM_INFO_COUNT = 16
};

/* Build the per-GPU context: internal DOF renumbering, band profile, element
 * colouring, device tables, workspace.  Replaces the setup the reference does
 * in fem_preprocess.py:291-443 + fem_solver_tf.py:378-396 (host side). */
int vbfem_create(vbfem_t **out, const vbfem_mesh *mesh, int device);

/* Options beyond the reference's default cards.  stype = section['stype'] (model_property_cards.py:28): 2 = plane
 * strain (what the TF path executes unconditionally, src/mat_subroutine_tf.py:54-56), 1 = plane stress (the NumPy
 * twin's other branch, src/mat_subroutine.py:283-290: Ce = E/(1-v^2) [[1,v,0],[v,1,0],[0,0,(1-v)/2]], sigma_zz = 0,
 * eps_33 = -v/(1-v)(eps_xx+eps_yy)).  Plane stress runs on the generic kernel (all modes incl. the adjoint).
 * Body forces (part['body'], src/mat_subroutine_tf.py:157-158) need no option: they are sample-independent and
 * enter through the load vector `pf` (see fem_solver.body_force_vector). */
typedef struct vbfem_options {
    int32_t stype;
    int32_t reserved[7];
} vbfem_options;
int vbfem_create_ex(vbfem_t **out, const vbfem_mesh *mesh, const vbfem_options *opt /* NULL: defaults */, int device);

/* The host-side plan vbfem_create would make for this mesh and observation set-up, WITHOUT touching
 * a GPU (unit tests of the numbering / orientation / front split): out[0] = kernel variant (4 = warp-per-sample
 * kernel, 2 = on-chip two-front kernel, 3 = blocked panel kernel, 0 = generic kernel), out[1] = order n, out[2] = half bandwidth, out[3] = first
 * middle row pT, out[4] = bottom-front columns nB, out[5] = 1 if the band order was reversed so that
 * it ends at the observed node, out[6] = shared memory per CTA in bytes (of the
 * forward / fused-adjoint launches), out[7] = generation of the warp kernel (1, 2; 0 for the other kernels).
 * smem_per_sm: shared memory per SM assumed for the fit test (<= 0: 233472, B200). */
int vbfem_plan(const vbfem_mesh *mesh, int64_t smem_per_sm, int64_t *out /* [8] */);
void vbfem_destroy(vbfem_t *h);
const char *vbfem_last_error(void);
int vbfem_info(const vbfem_t *h, int64_t *out /* [VBFEM_INFO_COUNT] */);

/* Size every per-sample buffer of the handle (status words, kept Jacobians, ELBO scratch, host staging)
 * for batches of up to n_samples_max.  Optional outside stream capture (the buffers grow on demand);
 * REQUIRED before capturing launches in a CUDA graph: a launch that would have to allocate during
 * capture fails with -6 instead. */
int vbfem_reserve(vbfem_t *h, int64_t n_samples_max);

/* y,h = fem_fh_fun_loop_rev(x): for each sample theta->(E,nu)
 * (data_generation_2sam_more_loss.py:181-186), assemble
 * (fem_solver_tf.py:229-341, mat_subroutine_tf.py:23-110), solve
 * (fem_solver_tf.py:129-153), observe u at obs_node and the von Mises stress
 * at (obs_ele, obs_gp) (fem_postprocess.py:172-185).  With keep_factor != 0
 * the library also keeps what vbfem_backward needs: the 4x2 Jacobian d(y, h)/dx
 * per sample, in a buffer owned by the handle (every such call gets a new ticket). */
int vbfem_forward(vbfem_t *h, int64_t n_samples, const double *x_dev /* [N][2] */,
                  double *y_dev /* [N][2] */, double *h_dev /* [N][2] */,
                  int keep_factor, void *stream);

/* gx = d(sum(gy*y) + sum(gh*h))/dx for the batch of the last
 * vbfem_forward(keep_factor=1): the discrete adjoint that tape.gradient
 * (main_custom_training.py:252-256) derives through the TF graph (J^T g with
 * the stored Jacobians, or an adjoint solve with the stored factor). */
int vbfem_backward(vbfem_t *h, int64_t n_samples, const double *gy_dev, const double *gh_dev,
                   double *gx_dev /* [N][2] */, void *stream);

/* Ticket of the Jacobians the handle currently keeps (0: none).  vbfem_backward_ticket applies them only
 * if `ticket` is still current and fails with -5 otherwise: a differentiable wrapper that took its
 * ticket right after vbfem_forward(keep_factor=1) can never consume another call's Jacobians. */
int64_t vbfem_keep_ticket(const vbfem_t *h);
int vbfem_backward_ticket(vbfem_t *h, int64_t ticket, int64_t n_samples, const double *gy_dev,
                          const double *gh_dev, double *gx_dev, void *stream);

/* Stateless pair for autograd nodes that own their state (the counterpart of one tf.custom_gradient
 * closure, main_custom_training.py:191-196 + 252-256): vbfem_forward_jac writes the per-sample
 * Jacobians [N][4][2] (rows y0, y1, h0, h1; columns x0, x1) into a CALLER-owned buffer,
 * vbfem_jac_vjp computes gx = J^T (gy, gh) from any such buffer.  Several forward calls may be
 * outstanding before their backward passes run. */
int vbfem_forward_jac(vbfem_t *h, int64_t n_samples, const double *x_dev, double *y_dev, double *h_dev,
                      double *jac_dev /* [N][8] */, void *stream);
int vbfem_jac_vjp(vbfem_t *h, int64_t n_samples, const double *jac_dev, const double *gy_dev,
                  const double *gh_dev, double *gx_dev, void *stream);

/* Fused forward + adjoint in one launch (no workspace round trip). */
int vbfem_forward_backward(vbfem_t *h, int64_t n_samples, const double *x_dev,
                           const double *gy_dev, const double *gh_dev,
                           double *y_dev, double *h_dev, double *gx_dev, void *stream);

/* Full fields for fem_test.py / fem_postprocess: u = sol_data['u_n1'] [N][ndof],
 * eps/sig = out_data['ele_strain'/'ele_stress'][:, :, :, 1] as [N][6][4][nele],
 * fint = sol_data['F_int'] [N][ndof]  (fem_solver_tf.py:310-341).
 * Any output pointer may be NULL.  If emat_dev != NULL it holds (E, nu) per
 * sample [N][2] and x_dev is ignored (fem_test.py uses the card values). */
int vbfem_fields(vbfem_t *h, int64_t n_samples, const double *x_dev, const double *emat_dev,
                 double *u_dev, double *sig_dev, double *eps_dev, double *fint_dev, void *stream);

/* Heterogeneous material: full fields (and the observations y, h) with ONE (E, nu) PER ELEMENT and sample,
 * emat_dev [N][nele][2] -- the generality the reference's cards only hint at (material per part,
 * src/mat_subroutine.py:24-25).  Forward only.  Any output pointer may be NULL. */
int vbfem_fields_elementwise(vbfem_t *h, int64_t n_samples, const double *emat_dev, double *y_dev, double *h_dev,
                             double *u_dev, double *sig_dev, double *eps_dev, double *fint_dev, void *stream);

/* Step-1 ELBO pieces (main_custom_training.py:183-235) for the flat sample
 * range [j_begin, j_end) of the B*S reparameterised samples
 * theta[b,s] = e[s]*sqrt(sig2[b]) + mu[b] (main_custom_training.py:199-209),
 * including the [B, B*S] broadcast of (y_point - f_data)
 * (main_custom_training.py:205,210-214).
 *   sums_dev[0..1] = sum_j f_j (per component), sums_dev[2] = sum_j |f_j|^2
 *   gmu_dev[B][2], gsig2_dev[B][2] = d(loss)/d(mu), d(loss)/d(sig2) through
 *   term2 only, restricted to the given sample range.
 * The caller (one rank per GPU) all-reduces these partials and adds the
 * closed-form term1/term3 parts. */
int vbfem_elbo_step1(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end,
                     const double *mu_dev, const double *sig2_dev, const double *e_dev,
                     const double *ybatch_dev, double sig_e, double *sums_dev, double *gmu_dev,
                     double *gsig2_dev, double *f_dev /* optional [j_end-j_begin][2] */, void *stream);

/* Step-2 ELBO data term (main_custom_training.py:338-364, term5): forward-only FEM over the flat
 * sample range [j_begin, j_end) of the B*S reparameterised samples of the FROZEN theta nets
 * (main_custom_training.py:305), reduced to the sufficient statistics of the [B, B*S] broadcast of
 * h_data against (z_mean, z_sig):
 *   sums_dev[0..1] = sum_j h_j (per component), sums_dev[2..3] = sum_j h_j^2.
 * The caller all-reduces them; term5 and its gradient w.r.t. the z nets are closed-form in these
 * sums (elbo.Step2Loss).  h_dev (optional) receives h [j_end-j_begin][2]. */
int vbfem_elbo_step2(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end,
                     const double *mu_dev, const double *sig2_dev, const double *e_dev,
                     double *sums_dev, double *h_dev /* optional */, void *stream);

/* ---- ELBO partial sums all-reduced over NVLink peer memory (one process per GPU) ----------------
 * The only exchange of the sharded ELBO step is the sum over ranks of 3 + 4B doubles
 * (step 1) or 4 doubles (step 2).  Instead of a collective AFTER the reduction kernel, the reduction
 * kernel stores its partial sums into a mailbox in every peer's memory (P2P stores), raises a
 * sequence flag there, waits for the peers' flags and adds the partials in rank order: one launch
 * less per step, no separate collective, bit-identical totals on every rank.
 *
 *   vbfem_peer_open      allocate this rank's mailbox for `cap` doubles per exchange; writes its CUDA IPC
 *                        handle (64 bytes) to ipc_handle_out and/or its address to mailbox_out
 *   (host side: all-gather the handles in rank order -- torch.distributed, MPI, a file ...)
 *   vbfem_peer_connect   map the peers' mailboxes: ipc_handles = world x 64 bytes (one process per
 *                        GPU), or mailboxes = world addresses (several handles of ONE process).
 *                        The caller puts a host barrier between connect and the first exchange.
 *   vbfem_peer_allreduce in-place sum of buf_dev[0, n) over the ranks (stand-alone form)
 *   vbfem_peer_status    synchronises; number of exchanges done so far, or -7 if a wait timed out
 *                        (a peer never arrived; VBFEM_PEER_TIMEOUT_MS, default 10000)
 * Every rank must issue the same sequence of exchanges (like any collective).  All calls are legal
 * inside CUDA-graph capture except open / connect / status. */
int vbfem_peer_open(vbfem_t *h, int32_t rank, int32_t world, int32_t cap_doubles,
                    void *ipc_handle_out /* 64 bytes or NULL */, void **mailbox_out /* or NULL */);
int vbfem_peer_connect(vbfem_t *h, const void *ipc_handles /* world x 64 bytes or NULL */,
                       void *const *mailboxes /* world pointers or NULL */);
int vbfem_peer_allreduce(vbfem_t *h, double *buf_dev, int32_t n, void *stream);
int64_t vbfem_peer_status(vbfem_t *h);

/* vbfem_elbo_step1 with the exchange fused into its reduction kernel:
 * totals_dev[3 + 4B] = [sums(3) | gmu(B x 2) | gsig2(B x 2)] summed over all ranks' sample ranges. */
int vbfem_elbo_step1_allreduce(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end,
                               const double *mu_dev, const double *sig2_dev, const double *e_dev,
                               const double *ybatch_dev, double sig_e, double *totals_dev,
                               double *f_dev /* optional */, void *stream);
/* The whole step-1 loss in the library: vbfem_elbo_step1 (+ the exchange when allreduce != 0, else the
 * range must be the whole step) followed by the closed-form KL terms (main_custom_training.py:183-185,
 * 226-235) and the loss value, so that a training step needs no framework arithmetic between the nets'
 * outputs and their gradients.  log_sig2_dev [B][2] is the log-variance net's output, sig2 = exp of it.
 *   out_dev[0]              loss = term1 - term2 - term3
 *   out_dev[1 .. 2B]        d loss / d mu            [B][2]
 *   out_dev[1+2B .. 4B]     d loss / d sig2          [B][2]
 *   out_dev[1+4B .. 6B]     d loss / d log_sig2      [B][2]  (direct dependence only; the caller chains sig2)
 * out_dev must hold 4 + 10B doubles (the totals are staged behind the results).  B <= 128. */
int vbfem_elbo_step1_loss(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end,
                          const double *mu_dev, const double *sig2_dev, const double *log_sig2_dev,
                          const double *e_dev, const double *ybatch_dev, double sig_e, int32_t allreduce,
                          double *out_dev, void *stream);

/* vbfem_elbo_step2 likewise: totals_dev[4] summed over all ranks. */
int vbfem_elbo_step2_allreduce(vbfem_t *h, int32_t B, int32_t S, int64_t j_begin, int64_t j_end,
                               const double *mu_dev, const double *sig2_dev, const double *e_dev,
                               double *totals_dev, double *h_dev /* optional */, void *stream);

/* Per-sample status words of the last launch (0 = ok, bit0 = non-positive or
 * non-finite pivot).  Synchronises.  Returns the number of flagged samples,
 * or a negative error. */
int64_t vbfem_status(vbfem_t *h, int32_t *flags_host /* [N] or NULL */, int64_t n_samples);

/* Host-buffer entry points (what a NumPy/TF caller uses): copy in from host
 * memory, run on the handle's own stream, copy out, synchronise.  Batches of up to 64
 * samples -- the one-sample-at-a-time callers, src/postprocess_lib.py:78-103 (Metropolis
 * log-posterior) -- skip the copies: the kernel works on mapped pinned host memory. */
int vbfem_forward_host(vbfem_t *h, int64_t n_samples, const double *x_host, double *y_host,
                       double *h_host);
int vbfem_forward_backward_host(vbfem_t *h, int64_t n_samples, const double *x_host,
                                const double *gy_host, const double *gh_host, double *y_host,
                                double *h_host, double *gx_host);

/* Test hook, GPU-free: the host tables of the blocked panel kernel for a mesh (see vbfem.cu). */
int64_t vbfem_debug_panel_tables(const vbfem_mesh *mesh, int which, void *out, int64_t cap_bytes);

/* Roofline denominators measured on this GPU: dependent-free DFMA loop
 * (TFLOP/s) and a device copy (GB/s, read+write bytes). */
int vbfem_measure_peaks(int device, double *fp64_tflops, double *copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* VBFEM_H */
