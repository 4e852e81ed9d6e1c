// Rendering-equation kernels: BRDF evaluation + hemisphere sum + tonemap.
//
// One warp per surface point; lanes stride the light directions so lvis[n, :] and
// spec[n, :] stream from HBM as coalesced 128-byte rows, light positions / areas /
// env-map texels sit in shared memory as float4, and the L-sum ends in a warp
// shuffle reduction.  Nothing of size [N, L, 3] is ever materialised (the reference
// builds four such tensors per env-map: nerfactor/models/nerfactor.py:325-342).
#include "nf_common.cuh"

namespace {

constexpr int WARPS = 8;
constexpr int E_CHUNK = 4;

struct PointCtx {  // per-point quantities (nerfactor.py:195-196, 212; microfacet.py:46-49)
  f3 pt, n1, v1, n2, v2, lambert;
  float alpha2_sq;   // (rough^2)^2 = alpha**2 with alpha = rough**2 (microfacet.py:54,102)
  float g_view;      // 2 / (1 + sqrt(1 + alpha^2 tan^2 theta_v))      (microfacet.py:83-89)
  float cos_v;       // n . v
  float abs_vn;
};

__device__ __forceinline__ PointCtx load_point(const nf_integrate_args& a, int i) {
  PointCtx c;
  c.pt = ld3(a.xyz_d + (size_t)i * 3);
  c.n1 = l2n(ld3(a.normal_d + (size_t)i * 3), 1e-6f);            // nerfactor.py:212
  c.v1 = l2n(ld3(a.cam_d + (size_t)i * 3) - c.pt, 1e-6f);         // shape.py:137-144
  f3 alb = ld3(a.albedo_d + (size_t)i * 3);
  c.lambert = mk3(alb.x / NF_PI_F, alb.y / NF_PI_F, alb.z / NF_PI_F);
  c.n2 = c.n1; c.v2 = c.v1; c.alpha2_sq = 0.f; c.g_view = 0.f; c.cos_v = 0.f; c.abs_vn = 0.f;
  if (a.brdf_kind == 0) {
    c.n2 = l2n(c.n1, 1e-6f);                                     // microfacet.py:47-49
    c.v2 = l2n(c.v1, 1e-6f);
    float rough = a.rough_d[i];
    float alpha = rough * rough;                                  // microfacet.py:54
    c.alpha2_sq = alpha * alpha;
    c.cos_v = dot3(c.n2, c.v2);                                   // microfacet.py:77
    float cv2 = fminf(fmaxf(c.cos_v * c.cos_v, 0.f), 1.f);        // :82-84
    float tan2 = fmaxf(divide_no_nan(1.f - cv2, cv2), 0.f);       // :85-87
    c.g_view = 2.f / (1.f + sqrtf(1.f + c.alpha2_sq * tan2));     // :88-89 (denominator >= 2)
    c.abs_vn = fabsf(c.cos_v);
  }
  return c;
}

// Returns the achromatic specular term and w = lvis*[cos>0]*cos*area for one pair.
//
// GGX lobe of brdf/microfacet/microfacet.py:30-111, algebraically reduced to one division:
//   D = a2 chi_d / (pi cm^4 (a2 + tan^2)^2)  with tan^2 = (1 - cm^2)/cm^2
//     = a2 chi_d / (pi u^2),  u = a2 cm^2 + 1 - cm^2           (cm^2 != 0; cm = 0 -> chi_d = 0)
//   G = chi_g * g_view (per-point, view-side only), chi_g = [(h.v)/(n.v) > 0] = [(h.v)(n.v) > 0]
//   spec = F G D / (4 |l.n| |v.n|) = F g_view a2 chi / (4 pi u^2 |l.n| |v.n|)   (0 if the
//   denominator is 0, tf.math.divide_no_nan).
// The reference re-normalises the already unit light direction inside Microfacet
// (microfacet.py:46); that second normalisation is the identity to 1 ulp and is skipped.
__device__ __forceinline__ void eval_pair(const nf_integrate_args& a, const PointCtx& c,
                                          float4 lx, float lvis, float spec_in,
                                          float& spec, float& w) {
  f3 d = mk3(lx.x, lx.y, lx.z) - c.pt;
  float inv = rsqrtf(fmaxf(dot3(d, d), 1e-6f));                  // shape.py:128-135
  f3 l1 = d * inv;
  float cosv = dot3(l1, c.n1);                                   // nerfactor.py:325
  w = (cosv > 0.f ? lvis : 0.f) * cosv * lx.w;                   // :329-335 (lx.w = area)
  if (a.brdf_kind == 0) {
    f3 hs = l1 + c.v2;
    f3 h = hs * rsqrtf(fmaxf(dot3(hs, hs), 1e-6f));              // microfacet.py:51-52
    float om = 1.f - dot3(l1, h);
    float om2 = om * om;
    float f = a.f0 + (1.f - a.f0) * (om2 * om2 * om);            // :106-111
    float cm = dot3(h, c.n2);                                    // :96
    float cm2 = cm * cm;
    float u = fmaf(c.alpha2_sq, cm2, 1.f - cm2);
    float hv = dot3(h, c.v2);                                    // :78
    bool on = (cm > 0.f) && (hv * c.cos_v > 0.f);                // chi_d, chi_g (:80-81, :97)
    float ldn = dot3(l1, c.n2);
    float den = (4.f * NF_PI_F) * (u * u) * (fabsf(ldn) * c.abs_vn);
    float num = f * (c.g_view * c.alpha2_sq);
    spec = (on && den != 0.f) ? __fdividef(num, den) : 0.f;      // :58-61
  } else {
    spec = spec_in * a.spec_scale;                               // nerfactor.py:460
  }
}

__device__ __forceinline__ float tonemap(float x, int srgb) {
  x = fminf(fmaxf(x, 0.f), 1.f);                                 // nerfactor.py:338
  return srgb ? linear2srgb_dev(x) : x;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// smem: lx4[L] (lxyz, area) ++ light4[ec][L] (texel of light idx(l), padded)
__global__ void __launch_bounds__(WARPS * 32) integrate_kernel(const nf_integrate_args a) {
  extern __shared__ float4 sm4[];
  float4* lx4 = sm4;
  float4* lt4 = sm4 + a.n_lights;
  const int L = a.n_lights;
  for (int l = threadIdx.x; l < L; l += blockDim.x)
    lx4[l] = make_float4(a.lxyz_d[l * 3], a.lxyz_d[l * 3 + 1], a.lxyz_d[l * 3 + 2], a.lareas_d[l]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e0 = 0; e0 < a.n_envmaps; e0 += E_CHUNK) {
    const int ec = min(E_CHUNK, a.n_envmaps - e0);
    __syncthreads();
    for (int i = threadIdx.x; i < ec * L; i += blockDim.x) {
      int e = i / L, l = i % L;
      int px = a.light_idx_d ? a.light_idx_d[l] : l;
      const float* t = a.light_d + ((size_t)(e0 + e) * a.envmap_pixels + px) * 3;
      lt4[e * L + l] = make_float4(t[0], t[1], t[2], 0.f);
    }
    __syncthreads();
    for (int i = blockIdx.x * WARPS + warp; i < a.n; i += gridDim.x * WARPS) {
      PointCtx c = load_point(a, i);
      float acc[E_CHUNK][3];
#pragma unroll
      for (int e = 0; e < E_CHUNK; ++e) acc[e][0] = acc[e][1] = acc[e][2] = 0.f;
      const float* lv = a.lvis_d + (size_t)i * L;
      const float* sp = a.brdf_kind == 1 ? a.spec_d + (size_t)i * L : nullptr;
      for (int l = lane; l < L; l += 32) {
        float spec, w;
        eval_pair(a, c, lx4[l], __ldg(lv + l), sp ? __ldg(sp + l) : 0.f, spec, w);
        float b0 = spec + c.lambert.x, b1 = spec + c.lambert.y, b2 = spec + c.lambert.z;
#pragma unroll
        for (int e = 0; e < E_CHUNK; ++e) {
          if (e < ec) {
            float4 t = lt4[e * L + l];
            // brdf * (lvis * light) * cos * area, nerfactor.py:334-336
            acc[e][0] += b0 * (w * t.x);
            acc[e][1] += b1 * (w * t.y);
            acc[e][2] += b2 * (w * t.z);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < E_CHUNK; ++e) {
        if (e < ec) {
          float r = warp_sum(acc[e][0]), g = warp_sum(acc[e][1]), b = warp_sum(acc[e][2]);
          if (lane == 0) {
            float* o = a.rgb_d + ((size_t)i * a.n_envmaps + e0 + e) * 3;
            o[0] = tonemap(r, a.linear2srgb);
            o[1] = tonemap(g, a.linear2srgb);
            o[2] = tonemap(b, a.linear2srgb);
          }
        }
      }
    }
  }
}

// OLAT: env-map l = inten * onehot(l) + ambient  ->  rgb[l] = ambient * sum_l' c[l'] + inten * c[l]
__global__ void __launch_bounds__(WARPS * 32)
integrate_olat_kernel(const nf_integrate_args a, float inten, float ambient, float* out) {
  extern __shared__ float4 sm4[];
  float4* lx4 = sm4;
  const int L = a.n_lights;
  for (int l = threadIdx.x; l < L; l += blockDim.x)
    lx4[l] = make_float4(a.lxyz_d[l * 3], a.lxyz_d[l * 3 + 1], a.lxyz_d[l * 3 + 2], a.lareas_d[l]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = blockIdx.x * WARPS + warp; i < a.n; i += gridDim.x * WARPS) {
    PointCtx c = load_point(a, i);
    const float* lv = a.lvis_d + (size_t)i * L;
    const float* sp = a.brdf_kind == 1 ? a.spec_d + (size_t)i * L : nullptr;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (ambient != 0.f) {
      for (int l = lane; l < L; l += 32) {
        float spec, w;
        eval_pair(a, c, lx4[l], __ldg(lv + l), sp ? __ldg(sp + l) : 0.f, spec, w);
        s0 += (spec + c.lambert.x) * (w * ambient);
        s1 += (spec + c.lambert.y) * (w * ambient);
        s2 += (spec + c.lambert.z) * (w * ambient);
      }
      s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    }
    for (int l = lane; l < L; l += 32) {
      float spec, w;
      eval_pair(a, c, lx4[l], __ldg(lv + l), sp ? __ldg(sp + l) : 0.f, spec, w);
      float* o = out + ((size_t)i * L + l) * 3;
      o[0] = tonemap(s0 + (spec + c.lambert.x) * (w * inten), a.linear2srgb);
      o[1] = tonemap(s1 + (spec + c.lambert.y) * (w * inten), a.linear2srgb);
      o[2] = tonemap(s2 + (spec + c.lambert.z) * (w * inten), a.linear2srgb);
    }
  }
}

int check_args(nf_ctx* ctx, const nf_integrate_args* a) {
  NF_CHECK_ARG(ctx, a, "null args");
  NF_CHECK_ARG(ctx, a->n >= 0 && a->n_lights > 0 && a->n_lights <= 4096, "bad n / n_lights");
  NF_CHECK_ARG(ctx, a->brdf_kind == 0 || a->brdf_kind == 1, "bad brdf_kind");
  if (a->n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, a->xyz_d && a->normal_d && a->cam_d && a->albedo_d && a->lvis_d &&
                        a->lxyz_d && a->lareas_d, "null buffer");
  NF_CHECK_ARG(ctx, a->brdf_kind == 0 ? a->rough_d != nullptr : a->spec_d != nullptr,
               "missing rough_d / spec_d");
  return NF_OK;
}

}  // namespace

extern "C" {

int nf_integrate_fwd(nf_ctx* ctx, const nf_integrate_args* a, void* stream) {
  int rc = check_args(ctx, a);
  if (rc != NF_OK) return rc;
  if (a->n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, a->n_envmaps >= 1 && a->light_d && a->rgb_d, "missing env-maps / output");
  NF_CHECK_ARG(ctx, a->envmap_pixels >= 1, "bad envmap_pixels");
  int ec = a->n_envmaps < E_CHUNK ? a->n_envmaps : E_CHUNK;
  size_t sm = sizeof(float4) * (size_t)a->n_lights * (1 + ec);
  NF_CHECK_ARG(ctx, sm <= ctx->smem_optin, "n_lights too large for shared memory");
  NF_CUDA(ctx, cudaFuncSetAttribute(integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int blocks_needed = (a->n + WARPS - 1) / WARPS;
  int grid = ctx->sm_count * 8;
  if (grid > blocks_needed) grid = blocks_needed;
  integrate_kernel<<<grid, WARPS * 32, sm, (cudaStream_t)stream>>>(*a);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_integrate_olat_fwd(nf_ctx* ctx, const nf_integrate_args* a, float olat_inten,
                          float ambient, float* rgb_olat_d, void* stream) {
  int rc = check_args(ctx, a);
  if (rc != NF_OK) return rc;
  if (a->n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, rgb_olat_d, "null output");
  NF_CHECK_ARG(ctx, a->light_idx_d == nullptr, "OLAT needs the identity light map");
  size_t sm = sizeof(float4) * (size_t)a->n_lights;
  NF_CUDA(ctx, cudaFuncSetAttribute(integrate_olat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  int blocks_needed = (a->n + WARPS - 1) / WARPS;
  int grid = ctx->sm_count * 8;
  if (grid > blocks_needed) grid = blocks_needed;
  integrate_olat_kernel<<<grid, WARPS * 32, sm, (cudaStream_t)stream>>>(*a, olat_inten, ambient, rgb_olat_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

}  // extern "C"
