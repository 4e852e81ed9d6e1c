// nf_stageB_fused_fwd: light-visibility network -> BRDF -> rendering equation behind one call
// (nerfactor/models/nerfactor.py:217-226, 262-266, 315-342), without the [N, L] light-visibility /
// BRDF tensors resident in HBM:
//   * default: the kernels the separate entry points launch, run over point chunks small enough
//     for a chunk's [c, L] rows to stay in the 126 MB L2 between producer and consumer
//     (workspace from the caller);
//   * NF_STAGEB_SINGLE=1 (microfacet BRDF, one env-map, L <= 512, tensor-core precision): ONE
//     kernel -- the rendering equation is evaluated in the head epilogue of the light-visibility
//     network (csrc/nf_mlp_tc.cu), the visibility values never leave the SM.
#include "nf_common.cuh"

int nf_tc_lvis_render_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, int n, float xyz_scale,
                             const float* lxyz, int L, const float* normal, const float* cam,
                             const float* albedo, const float* rough, const float* lareas,
                             const float* light, const int* light_idx, float f0, int srgb,
                             float* lvis, float* rgb, int precision, cudaStream_t st);

int nf_tc_lvis_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, int n, float xyz_scale,
                      const float* lxyz, int L, float* lvis, int precision, cudaStream_t st,
                      const float* xyz_dir, const float* cull_normal);

namespace {
// points per chunk of the chunked path: [c, L] fp32 rows of <= 48 MB (two of them for the learned
// BRDF: both stay inside the 126 MB L2 between producer and consumer)
int chunk_points(int n, int L) {
  long long c = (48ll << 20) / ((long long)L * 4);
  c = c / 256 * 256;
  if (c < 256) c = 256;
  return (int)(c < n ? c : n);
}
size_t align256(size_t x) { return (x + 255) / 256 * 256; }
// The one-kernel variant (rendering equation in the head epilogue of the visibility network) is
// correct and keeps the visibility values on the SM, but measured SLOWER than the chunked pair
// (38.0 vs 32.2 + 0.9 ms at 640 k x 512: the epilogue -> MMA hand-over chain of that kernel is its
// critical path, profiles/r2_k2_analysis.md), so it is opt-in: NF_STAGEB_SINGLE=1.
bool single_kernel(const nf_stageb_args* a, int precision) {
  static const bool want = [] { const char* e = getenv("NF_STAGEB_SINGLE"); return e && e[0] == '1'; }();
  return want && a->brdf_kind == 0 && a->n_envmaps == 1 && a->n_lights <= 512 &&
         (precision == NF_PREC_F16 || precision == NF_PREC_BF16);
}
}  // namespace

extern "C" {

size_t nf_stageB_fused_workspace_bytes(const nf_stageb_args* a, int precision) {
  if (!a || a->n <= 0 || single_kernel(a, precision)) return 0;
  const size_t c = (size_t)chunk_points(a->n, a->n_lights);
  const size_t rows = align256(c * a->n_lights * 4);
  return (a->lvis_d ? 0 : rows) + (a->brdf_kind == 1 ? rows : 0);
}

int nf_stageB_fused_fwd(nf_ctx* ctx, const nf_mlp* mlp_lvis, const nf_mlp* mlp_brdf,
                        const nf_stageb_args* a, int precision, void* workspace_d,
                        size_t workspace_bytes, void* stream) {
  NF_CHECK_ARG(ctx, a && mlp_lvis, "null argument");
  NF_CHECK_ARG(ctx, a->n >= 0 && a->n_lights > 0 && a->n_envmaps >= 1, "bad sizes");
  NF_CHECK_ARG(ctx, a->brdf_kind == 0 || a->brdf_kind == 1, "bad brdf_kind");
  if (a->n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, a->xyz_d && a->normal_d && a->cam_d && a->albedo_d && a->lxyz_d && a->lareas_d &&
                        a->light_d && a->rgb_d, "null buffer");
  NF_CHECK_ARG(ctx, a->brdf_kind == 0 ? a->rough_d != nullptr : (a->z_d != nullptr && mlp_brdf != nullptr),
               "missing rough_d / (z_d, mlp_brdf)");
  NF_CHECK_ARG(ctx, mlp_lvis->d.kind == NF_MLP_LVIS && mlp_lvis->d.out_dim == 1,
               "mlp_lvis is not NF_MLP_LVIS");
  cudaStream_t st = (cudaStream_t)stream;
  if (single_kernel(a, precision))
    return nf_tc_lvis_render_launch(ctx, mlp_lvis, a->xyz_d, a->n, a->xyz_scale, a->lxyz_d, a->n_lights,
                                    a->normal_d, a->cam_d, a->albedo_d, a->rough_d, a->lareas_d,
                                    a->light_d, a->light_idx_d, a->f0, a->linear2srgb, a->lvis_d,
                                    a->rgb_d, precision, st);
  NF_CHECK_ARG(ctx, workspace_bytes >= nf_stageB_fused_workspace_bytes(a, precision) &&
                        (workspace_d || workspace_bytes == 0),
               "workspace too small (nf_stageB_fused_workspace_bytes)");
  const int L = a->n_lights;
  const int c_max = chunk_points(a->n, L);
  uint8_t* w = (uint8_t*)workspace_d;
  float* lvis_ws = nullptr;
  if (!a->lvis_d) { lvis_ws = (float*)w; w += align256((size_t)c_max * L * 4); }
  float* spec_ws = a->brdf_kind == 1 ? (float*)w : nullptr;
  for (int p0 = 0; p0 < a->n; p0 += c_max) {
    const int c = a->n - p0 < c_max ? a->n - p0 : c_max;
    float* lv = a->lvis_d ? a->lvis_d + (size_t)p0 * L : lvis_ws;
    // the visibility tensor is not an output: only the lights the renderer does not multiply by
    // zero (nerfactor.py:329-330) go through the network (see nf_stageb_args.lvis_all_lights)
    const bool front_lit_only = (a->lvis_all_lights == 2 || (!a->lvis_d && a->lvis_all_lights == 0)) &&
                                (precision == NF_PREC_F16 || precision == NF_PREC_BF16);
    int rc = front_lit_only
                 ? nf_tc_lvis_launch(ctx, mlp_lvis, a->xyz_d + (size_t)p0 * 3, c, a->xyz_scale,
                                     a->lxyz_d, L, lv, precision, st, nullptr,
                                     a->normal_d + (size_t)p0 * 3)
                 : nf_lvis_fwd(ctx, mlp_lvis, a->xyz_d + (size_t)p0 * 3, c, a->xyz_scale, a->lxyz_d,
                               L, lv, precision, stream);
    if (rc != NF_OK) return rc;
    if (a->brdf_kind == 1) {
      rc = nf_brdf_learned_fwd(ctx, mlp_brdf, a->xyz_d + (size_t)p0 * 3, a->normal_d + (size_t)p0 * 3,
                               a->cam_d + (size_t)p0 * 3, a->z_d + (size_t)p0 * a->z_dim, c, a->lxyz_d,
                               L, spec_ws, precision, stream);
      if (rc != NF_OK) return rc;
    }
    nf_integrate_args ia;
    memset(&ia, 0, sizeof(ia));
    ia.n = c; ia.n_lights = L; ia.n_envmaps = a->n_envmaps; ia.envmap_pixels = a->envmap_pixels;
    ia.brdf_kind = a->brdf_kind; ia.linear2srgb = a->linear2srgb; ia.f0 = a->f0;
    ia.spec_scale = a->spec_scale;
    ia.xyz_d = a->xyz_d + (size_t)p0 * 3; ia.normal_d = a->normal_d + (size_t)p0 * 3;
    ia.cam_d = a->cam_d + (size_t)p0 * 3; ia.albedo_d = a->albedo_d + (size_t)p0 * 3;
    ia.rough_d = a->rough_d ? a->rough_d + p0 : nullptr;
    ia.spec_d = spec_ws; ia.lvis_d = lv; ia.lxyz_d = a->lxyz_d; ia.lareas_d = a->lareas_d;
    ia.light_d = a->light_d; ia.light_idx_d = a->light_idx_d;
    ia.rgb_d = a->rgb_d + (size_t)p0 * a->n_envmaps * 3;
    rc = nf_integrate_fwd(ctx, &ia, stream);
    if (rc != NF_OK) return rc;
  }
  return NF_OK;
}

}  // extern "C"
