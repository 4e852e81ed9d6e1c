// Whole ray-march operations behind one C-ABI call each (SURVEY.md 8b minimum exports):
//   nf_raymarch_depth_normal_fwd   geometry_from_nerf.compute_depth_and_normal  (gfn.py:249-319)
//   nf_raymarch_lvis_fwd           geometry_from_nerf.compute_light_visibility  (gfn.py:177-246)
// Each runs the stage kernels of this library back to back on the caller's stream over ray chunks,
// with every intermediate ([n, S] depths / densities / weights / per-sample normals, the compacted
// front-lit ray list) in a caller-provided workspace: nothing is allocated behind the caller, and
// the light-visibility march does its front-lit test, compaction, both marches and the
// 1 - sum(w) scatter on the device (the Python layer used torch.nonzero / index_select before).
#include "nf_common.cuh"

namespace {

constexpr int DN_CHUNK = 32768;        // rays per pass of the camera march
constexpr int LV_CHUNK = 1 << 19;      // (point, light) pairs per pass of the light march

// Front-lit test of gfn.py:196-215 for the pairs [g0, g0 + count) and compaction of the front-lit
// ones: ray origin / direction and the pair index, slots handed out by a warp-aggregated atomic
// (the order inside a chunk is irrelevant: every pair is marched independently and scattered
// back by its index).
__global__ void __launch_bounds__(256)
lvis_compact_kernel(const float* __restrict__ surf, const float* __restrict__ normal,
                    const float* __restrict__ lxyz, int L, long long g0, int count,
                    float* __restrict__ rayo, float* __restrict__ rayd,
                    int* __restrict__ pair, int* __restrict__ counter) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool lit = false;
  f3 s = mk3(0.f, 0.f, 0.f), d = mk3(0.f, 0.f, 0.f);
  if (j < count) {
    const long long g = g0 + j;
    const long long i = g / L;
    const int l = (int)(g % L);
    s = ld3(surf + i * 3);
    d = l2n(ld3(lxyz + l * 3) - s, 1e-12f);                       // gfn.py:197-198
    lit = dot3(d, ld3(normal + i * 3)) > 0.f;                     // gfn.py:205-206
  }
  const unsigned m = __ballot_sync(0xffffffffu, lit);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
  if (lit) {
    const int k = base + __popc(m & ((1u << lane) - 1u));
    rayo[k * 3] = s.x; rayo[k * 3 + 1] = s.y; rayo[k * 3 + 2] = s.z;
    rayd[k * 3] = d.x; rayd[k * 3 + 1] = d.y; rayd[k * 3 + 2] = d.z;
    pair[k] = j;
  }
}

// lvis[g0 + pair[k]] = 1 - occu[k]   (gfn.py:240-244)
__global__ void lvis_scatter_kernel(const float* __restrict__ occu, const int* __restrict__ pair,
                                    int k_count, long long g0, float* __restrict__ lvis) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < k_count) lvis[g0 + pair[k]] = 1.f - occu[k];
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" {

size_t nf_raymarch_depth_normal_workspace_bytes(int n_rays, int n_coarse, int n_fine) {
  const size_t c = (size_t)(n_rays < DN_CHUNK ? (n_rays > 0 ? n_rays : 1) : DN_CHUNK);
  const size_t S = (size_t)n_coarse + n_fine;
  // z_c, sigma_c, w_c [c, Sc];  z_all, sigma_f [c, S];  normal [c, S, 3]
  return 3 * align256(c * n_coarse * 4) + 2 * align256(c * S * 4) + align256(c * S * 12);
}

int nf_raymarch_depth_normal_fwd(nf_ctx* ctx, const nf_mlp* mlp_coarse, const nf_mlp* mlp_fine,
                                 const float* rayo_d, const float* rayd_d, int n_rays, float near,
                                 float far, int n_coarse, int n_fine, int lin_in_disp,
                                 const float* bbox_host, int precision, void* workspace_d,
                                 size_t workspace_bytes, float* occu_d, float* depth_d,
                                 float* normal_d, void* stream) {
  NF_CHECK_ARG(ctx, mlp_coarse && mlp_fine, "null network");
  NF_CHECK_ARG(ctx, n_rays >= 0 && n_coarse >= 3 && n_fine >= 1, "bad sizes");
  if (n_rays == 0) return NF_OK;
  NF_CHECK_ARG(ctx, rayo_d && rayd_d && occu_d && depth_d && normal_d && workspace_d, "null buffer");
  NF_CHECK_ARG(ctx, workspace_bytes >= nf_raymarch_depth_normal_workspace_bytes(n_rays, n_coarse, n_fine),
               "workspace too small (nf_raymarch_depth_normal_workspace_bytes)");
  const int c_max = n_rays < DN_CHUNK ? n_rays : DN_CHUNK;
  const int S = n_coarse + n_fine;
  uint8_t* w = (uint8_t*)workspace_d;
  float* z_c = (float*)w;       w += align256((size_t)c_max * n_coarse * 4);
  float* sg_c = (float*)w;      w += align256((size_t)c_max * n_coarse * 4);
  float* w_c = (float*)w;       w += align256((size_t)c_max * n_coarse * 4);
  float* z_all = (float*)w;     w += align256((size_t)c_max * S * 4);
  float* sg_f = (float*)w;      w += align256((size_t)c_max * S * 4);
  float* nrm = (float*)w;
  // the forward + input-gradient kernel has no split-encoding variant
  const int prec_n = precision == NF_PREC_F16E ? NF_PREC_F16 : precision;
  for (int r0 = 0; r0 < n_rays; r0 += c_max) {
    const int c = n_rays - r0 < c_max ? n_rays - r0 : c_max;
    const float* ro = rayo_d + (size_t)r0 * 3;
    const float* rd = rayd_d + (size_t)r0 * 3;
    int rc;
    if ((rc = nf_gen_z(ctx, near, far, n_coarse, c, lin_in_disp, nullptr, z_c, stream))) return rc;
    if ((rc = nf_sigma_fwd(ctx, mlp_coarse, ro, rd, z_c, c, n_coarse, bbox_host, sg_c, precision, stream))) return rc;
    if ((rc = nf_composite(ctx, sg_c, z_c, ro, rd, nullptr, c, n_coarse, w_c, occu_d + r0, depth_d + r0,
                           nullptr, nullptr, stream))) return rc;
    if ((rc = nf_gen_z_fine(ctx, z_c, w_c, c, n_coarse, n_fine, z_all, stream))) return rc;
    if ((rc = nf_sigma_normal_fwd(ctx, mlp_fine, ro, rd, z_all, c, S, bbox_host, sg_f, nrm, prec_n, stream))) return rc;
    if ((rc = nf_composite(ctx, sg_f, z_all, ro, rd, nrm, c, S, nullptr, occu_d + r0, depth_d + r0,
                           nullptr, normal_d + (size_t)r0 * 3, stream))) return rc;
  }
  return NF_OK;
}

size_t nf_raymarch_lvis_workspace_bytes(int n_pts, int n_lights, int n_coarse, int n_fine) {
  long long pairs = (long long)n_pts * n_lights;
  const size_t c = (size_t)(pairs < LV_CHUNK ? (pairs > 0 ? pairs : 1) : LV_CHUNK);
  const size_t S = (size_t)n_coarse + n_fine;
  // rayo, rayd [c, 3]; pair [c]; counter; z_c, sigma_c, w_c [c, Sc]; z_all, sigma_f [c, S]; occu, depth [c]
  return 2 * align256(c * 12) + align256(c * 4) + 256 + 3 * align256(c * n_coarse * 4) +
         2 * align256(c * S * 4) + 2 * align256(c * 4);
}

int nf_raymarch_lvis_fwd(nf_ctx* ctx, const nf_mlp* mlp_coarse, const nf_mlp* mlp_fine,
                         const float* surf_d, const float* normal_d, int n_pts,
                         const float* lxyz_d, int n_lights, float lvis_near, float lvis_far,
                         int n_coarse, int n_fine, int lin_in_disp, const float* bbox_host,
                         int precision, void* workspace_d, size_t workspace_bytes, float* lvis_d,
                         void* stream) {
  NF_CHECK_ARG(ctx, mlp_coarse && mlp_fine, "null network");
  NF_CHECK_ARG(ctx, n_pts >= 0 && n_lights > 0 && n_coarse >= 3 && n_fine >= 1, "bad sizes");
  const long long pairs = (long long)n_pts * n_lights;
  if (pairs == 0) return NF_OK;
  NF_CHECK_ARG(ctx, surf_d && normal_d && lxyz_d && lvis_d && workspace_d, "null buffer");
  NF_CHECK_ARG(ctx, workspace_bytes >= nf_raymarch_lvis_workspace_bytes(n_pts, n_lights, n_coarse, n_fine),
               "workspace too small (nf_raymarch_lvis_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  const int c_max = (int)(pairs < LV_CHUNK ? pairs : LV_CHUNK);
  const int S = n_coarse + n_fine;
  uint8_t* w = (uint8_t*)workspace_d;
  float* ro = (float*)w;        w += align256((size_t)c_max * 12);
  float* rd = (float*)w;        w += align256((size_t)c_max * 12);
  int* pair = (int*)w;          w += align256((size_t)c_max * 4);
  int* counter = (int*)w;       w += 256;
  float* z_c = (float*)w;       w += align256((size_t)c_max * n_coarse * 4);
  float* sg_c = (float*)w;      w += align256((size_t)c_max * n_coarse * 4);
  float* w_c = (float*)w;       w += align256((size_t)c_max * n_coarse * 4);
  float* z_all = (float*)w;     w += align256((size_t)c_max * S * 4);
  float* sg_f = (float*)w;      w += align256((size_t)c_max * S * 4);
  float* occu = (float*)w;      w += align256((size_t)c_max * 4);
  float* depth = (float*)w;
  NF_CUDA(ctx, cudaMemsetAsync(lvis_d, 0, (size_t)pairs * 4, st));          // back-lit pairs stay 0
  for (long long g0 = 0; g0 < pairs; g0 += c_max) {
    const int c = (int)(pairs - g0 < c_max ? pairs - g0 : c_max);
    NF_CUDA(ctx, cudaMemsetAsync(counter, 0, 4, st));
    lvis_compact_kernel<<<(c + 255) / 256, 256, 0, st>>>(surf_d, normal_d, lxyz_d, n_lights, g0, c,
                                                          ro, rd, pair, counter);
    NF_LAUNCH_CHECK(ctx);
    // The number of front-lit pairs sizes the following launches: one 4-byte read-back and a
    // stream synchronisation per chunk (the reference's tf.boolean_mask does the same, gfn.py:210).
    int k = 0;
    NF_CUDA(ctx, cudaMemcpyAsync(&k, counter, 4, cudaMemcpyDeviceToHost, st));
    NF_CUDA(ctx, cudaStreamSynchronize(st));
    if (k == 0) continue;
    int rc;
    if ((rc = nf_gen_z(ctx, lvis_near, lvis_far, n_coarse, k, lin_in_disp, nullptr, z_c, stream))) return rc;
    if ((rc = nf_sigma_fwd(ctx, mlp_coarse, ro, rd, z_c, k, n_coarse, bbox_host, sg_c, precision, stream))) return rc;
    if ((rc = nf_composite(ctx, sg_c, z_c, ro, rd, nullptr, k, n_coarse, w_c, occu, depth, nullptr,
                           nullptr, stream))) return rc;
    if ((rc = nf_gen_z_fine(ctx, z_c, w_c, k, n_coarse, n_fine, z_all, stream))) return rc;
    if ((rc = nf_sigma_fwd(ctx, mlp_fine, ro, rd, z_all, k, S, bbox_host, sg_f, precision, stream))) return rc;
    if ((rc = nf_composite(ctx, sg_f, z_all, ro, rd, nullptr, k, S, nullptr, occu, depth, nullptr,
                           nullptr, stream))) return rc;
    lvis_scatter_kernel<<<(k + 255) / 256, 256, 0, st>>>(occu, pair, k, g0, lvis_d);
    NF_LAUNCH_CHECK(ctx);
  }
  return NF_OK;
}

}  // extern "C"
