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
constexpr int E_CHUNK = 8;   // env-maps per pass over the lvis rows (instantiated: 1 2 3 4 6 8)

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

// ---- packed-FP32 (FFMA2 / FMUL2 / FADD2, sm_100) version of the pair evaluation ------------
// Two consecutive lights per lane, every vector quantity as a float2 (light 2j, light 2j+1).
// Same formulas as eval_pair with these identities applied (each exact up to 1 ulp; l, v, n unit):
//   h . v = l . h = |l + v| / 2  (>= 0)                so chi_g = [(h.v)(n.v) > 0] = [n.v > 0]: per
//                                                      point, folded into K
//   h . n = (l.n + v.n) / |l + v|                      > 0 whenever l.n > 0 and n.v > 0: chi_d is
//                                                      implied by the front-lit test
//   l . n2 = l . n1 = cos                              (n2 = l2n(n1) is n1)
//   spec * cos / |l.n| = spec' for cos > 0: the division by |l.n| cancels against the cosine of
//   the rendering equation, so   brdf * w = (lambert * cos + F K / u^2) * [cos>0] lvis area,
//   K = [n.v > 0] g_view a2 / (4 pi |v.n|)  (0 when v.n = 0: tf.math.divide_no_nan,
//   microfacet.py:58-61),  u = 1 + (a2 - 1) (h.n)^2.
// NOT applied: |l + v|^2 = 2 + 2 l.v.  At grazing mirror configurations (|l + v| = 2 n.v -> 0) it
// cancels catastrophically -- measured 8e-5 (rough 0.4) / 1.6e-2 (rough 0.2) rel-L2 against an
// fp64 evaluation, vs 1e-6 / 2e-4 with the half vector formed component-wise as below
// (tools/diag_integrate.py).
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }
__device__ __forceinline__ float rsq_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_fast(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 dot3_2(float2 ax, float2 ay, float2 az, float2 bx, float2 by, float2 bz) {
  return __ffma2_rn(az, bz, __ffma2_rn(ay, by, __fmul2_rn(ax, bx)));
}

struct PointCtx2 {
  float2 npx, npy, npz;        // -point
  float2 nx, ny, nz;           // unit normal
  float2 vx, vy, vz;           // unit view direction
  float2 a2;                   // alpha^2 (= rough^4)
  float2 k;                    // [n.v > 0] g_view a2 / (4 pi |v.n|)
  float2 f0, omf0;
  float2 lr, lg, lb;           // albedo / pi
  float2 cos_v;                // n . v
};

// sw = specular * [cos>0] lvis area (already multiplied by the cosine, see above), dw = diffuse
// weight [cos>0] lvis area cos
template <int KIND>
__device__ __forceinline__ void eval_pair2(const nf_integrate_args& a, const PointCtx2& c,
                                           float2 lx, float2 ly, float2 lz, float2 la,
                                           float2 lvis, float2 spec_in, float2& sw, float2& dw) {
  const float2 dx = __fadd2_rn(lx, c.npx), dy = __fadd2_rn(ly, c.npy), dz = __fadd2_rn(lz, c.npz);
  const float2 dd = dot3_2(dx, dy, dz, dx, dy, dz);
  const float2 inv = make_float2(rsq_fast(fmaxf(dd.x, 1e-6f)), rsq_fast(fmaxf(dd.y, 1e-6f)));   // shape.py:128-135
  const float2 l1x = __fmul2_rn(dx, inv), l1y = __fmul2_rn(dy, inv), l1z = __fmul2_rn(dz, inv);
  const float2 cosv = dot3_2(l1x, l1y, l1z, c.nx, c.ny, c.nz);          // nerfactor.py:325
  const float2 wl = __fmul2_rn(lvis, la);                               // lvis * area
  dw = __fmul2_rn(wl, make_float2(fmaxf(cosv.x, 0.f), fmaxf(cosv.y, 0.f)));   // :329-335
  if (KIND == 0) {
    const float2 hx = __fadd2_rn(l1x, c.vx), hy = __fadd2_rn(l1y, c.vy), hz = __fadd2_rn(l1z, c.vz);
    const float2 hh = dot3_2(hx, hy, hz, hx, hy, hz);                           // |l + v|^2
    const float2 invh = make_float2(rsq_fast(fmaxf(hh.x, 1e-6f)), rsq_fast(fmaxf(hh.y, 1e-6f)));   // microfacet.py:51-52
    const float2 om = __ffma2_rn(__fmul2_rn(hh, invh), bc2(-0.5f), bc2(1.f));   // 1 - l.h
    const float2 om2 = __fmul2_rn(om, om);
    const float2 om5 = __fmul2_rn(__fmul2_rn(om2, om2), om);
    const float2 f = __ffma2_rn(om5, c.omf0, c.f0);                             // :106-111
    // u = a2 cm^2 + (1 - cm^2) with cm = (l + v).n / |l + v|.  Near the specular peak (h -> n)
    // 1 - cm^2 is a difference of two numbers ~1 whose rounding (6e-8) is comparable with a2
    // itself at low roughness (a2 = rough^4 = 1.6e-3 at 0.2): formed instead from the part of
    // l + v orthogonal to n,  u |l + v|^2 = a2 ((l + v).n)^2 + |(l + v) - ((l + v).n) n|^2.
    const float2 hn = __fadd2_rn(cosv, c.cos_v);                                // (l + v) . n, :96
    const float2 nhn = __fmul2_rn(hn, bc2(-1.f));
    const float2 tx = __ffma2_rn(nhn, c.nx, hx), ty = __ffma2_rn(nhn, c.ny, hy),
                 tz = __ffma2_rn(nhn, c.nz, hz);
    const float2 tt = dot3_2(tx, ty, tz, tx, ty, tz);
    const float2 U = __ffma2_rn(c.a2, __fmul2_rn(hn, hn), tt);                  // u |l + v|^2
    const float2 hU = __fmul2_rn(hh, make_float2(rcp_fast(U.x), rcp_fast(U.y)));  // 1 / u
    const float2 r = __fmul2_rn(hU, hU);
    const float2 u = U;
    const float2 sp = __fmul2_rn(__fmul2_rn(f, c.k), __fmul2_rn(wl, r));
    // front-lit (which implies chi_d once n.v > 0, folded into K) and a non-degenerate lobe
    sw = make_float2(fminf(cosv.x, u.x) > 0.f ? sp.x : 0.f, fminf(cosv.y, u.y) > 0.f ? sp.y : 0.f);
  } else {
    // learned BRDF: spec * learned_brdf_scale (nerfactor.py:460), times the diffuse weight
    sw = __fmul2_rn(__fmul2_rn(spec_in, bc2(a.spec_scale)), dw);
  }
}

// smem (floats, Lp = L rounded up to 2): X[Lp] Y[Lp] Z[Lp] AREA[Lp], then per env-map of the
// chunk R[Lp] G[Lp] B[Lp] (texel of light idx(l)); pad entries have area 0.
// EC = env-maps per pass (compile time: the accumulation loop has no branches).  The lvis /
// spec row of a point is fetched in one burst of up to 8 float2 per lane (512 lights) BEFORE the
// arithmetic, so each warp pays the DRAM latency once per point instead of once per 64 lights.
template <int EC, int KIND>
__global__ void __launch_bounds__(WARPS * 32, KIND == 0 && EC == 1 ? 3 : 1) integrate_kernel(const nf_integrate_args a) {
  extern __shared__ float4 sm4[];
  float* smf = reinterpret_cast<float*>(sm4);
  const int L = a.n_lights, Lp = (L + 1) & ~1;
  float* sx = smf; float* sy = sx + Lp; float* sz = sy + Lp; float* sa = sz + Lp;
  float* st = sa + Lp;
  for (int l = threadIdx.x; l < Lp; l += blockDim.x) {
    const bool in = l < L;
    sx[l] = in ? a.lxyz_d[l * 3] : 0.f;
    sy[l] = in ? a.lxyz_d[l * 3 + 1] : 0.f;
    sz[l] = in ? a.lxyz_d[l * 3 + 2] : 1.f;
    sa[l] = in ? a.lareas_d[l] : 0.f;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool vec = (L & 1) == 0;                 // rows of lvis / spec are 8-byte aligned
  for (int e0 = 0; e0 < a.n_envmaps; e0 += EC) {
    const int ec = min(EC, a.n_envmaps - e0);    // < EC only in a tail pass: extra maps read as 0
    __syncthreads();
    for (int i = threadIdx.x; i < EC * Lp; i += blockDim.x) {
      const int e = i / Lp, l = i % Lp;
      float r = 0.f, g = 0.f, b = 0.f;
      if (l < L && e < ec) {
        const int px = a.light_idx_d ? a.light_idx_d[l] : l;
        const float* t = a.light_d + ((size_t)(e0 + e) * a.envmap_pixels + px) * 3;
        r = t[0]; g = t[1]; b = t[2];
      }
      st[(e * 3 + 0) * Lp + l] = r;
      st[(e * 3 + 1) * Lp + l] = g;
      st[(e * 3 + 2) * Lp + l] = b;
    }
    __syncthreads();
    // A warp takes 32 consecutive points at a time: the per-point set-up (normalisations, the
    // view-side shadowing term) and the tone-mapping epilogue run lane-parallel, one point per
    // lane; inside, the points are visited one by one with all lanes striding the lights, the
    // point's constants arriving by shuffle from its lane.
    for (int base = (blockIdx.x * WARPS + warp) * 32; base < a.n; base += gridDim.x * WARPS * 32) {
      const int pi = base + lane;
      const bool pv = pi < a.n;
      const PointCtx s = load_point(a, pv ? pi : a.n - 1);
      // chi_g = [(h.v)(n.v) > 0] with h.v >= 0: a per-point gate on n.v (h.v = 0 needs l = -v,
      // where the pair is back-lit or the view is)
      const float my_k = s.cos_v > 0.f ? s.g_view * s.alpha2_sq / (4.f * NF_PI_F * s.abs_vn) : 0.f;
      float outv[EC][3];
#pragma unroll
      for (int e = 0; e < EC; ++e) outv[e][0] = outv[e][1] = outv[e][2] = 0.f;
      const int cnt = min(32, a.n - base);
      for (int j = 0; j < cnt; ++j) {
        const int i = base + j;
        const float* lv = a.lvis_d + (size_t)i * L;
        const float* sp = KIND == 1 ? a.spec_d + (size_t)i * L : nullptr;
        float2 acc[EC][3];
#pragma unroll
        for (int e = 0; e < EC; ++e) acc[e][0] = acc[e][1] = acc[e][2] = bc2(0.f);
        PointCtx2 c;
#define NF_BC(x) __shfl_sync(0xffffffffu, (x), j)
        c.npx = bc2(-NF_BC(s.pt.x)); c.npy = bc2(-NF_BC(s.pt.y)); c.npz = bc2(-NF_BC(s.pt.z));
        c.nx = bc2(NF_BC(s.n1.x)); c.ny = bc2(NF_BC(s.n1.y)); c.nz = bc2(NF_BC(s.n1.z));
        c.vx = bc2(NF_BC(s.v2.x)); c.vy = bc2(NF_BC(s.v2.y)); c.vz = bc2(NF_BC(s.v2.z));
        c.a2 = bc2(NF_BC(s.alpha2_sq));
        c.k = bc2(NF_BC(my_k));
        c.f0 = bc2(a.f0); c.omf0 = bc2(1.f - a.f0);
        c.lr = bc2(NF_BC(s.lambert.x)); c.lg = bc2(NF_BC(s.lambert.y)); c.lb = bc2(NF_BC(s.lambert.z));
        c.cos_v = bc2(NF_BC(s.cos_v));
#undef NF_BC
        for (int l0 = 0; l0 < L; l0 += 512) {
          float2 lvr[8], spr[KIND == 1 ? 8 : 1];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int l = l0 + 64 * q + 2 * lane;
            lvr[q] = bc2(0.f);
            if (KIND == 1) spr[q] = bc2(0.f);
            if (l < L) {
              if (vec) {
                lvr[q] = __ldg(reinterpret_cast<const float2*>(lv + l));
                if (KIND == 1) spr[q] = __ldg(reinterpret_cast<const float2*>(sp + l));
              } else {
                lvr[q] = make_float2(__ldg(lv + l), l + 1 < L ? __ldg(lv + l + 1) : 0.f);
                if (KIND == 1) spr[q] = make_float2(__ldg(sp + l), l + 1 < L ? __ldg(sp + l + 1) : 0.f);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int l = l0 + 64 * q + 2 * lane;
            if (l < L) {
              float2 sw, dw;
              eval_pair2<KIND>(a, c, *reinterpret_cast<const float2*>(sx + l),
                         *reinterpret_cast<const float2*>(sy + l), *reinterpret_cast<const float2*>(sz + l),
                         *reinterpret_cast<const float2*>(sa + l), lvr[q], spr[KIND == 1 ? q : 0], sw, dw);
              // brdf * (lvis * light) * cos * area (nerfactor.py:334-336) = (lambert dw + sw) * texel
              const float2 b0 = __ffma2_rn(c.lr, dw, sw), b1 = __ffma2_rn(c.lg, dw, sw),
                           b2 = __ffma2_rn(c.lb, dw, sw);
#pragma unroll
              for (int e = 0; e < EC; ++e) {
                const float* te = st + (size_t)e * 3 * Lp + l;
                acc[e][0] = __ffma2_rn(b0, *reinterpret_cast<const float2*>(te), acc[e][0]);
                acc[e][1] = __ffma2_rn(b1, *reinterpret_cast<const float2*>(te + Lp), acc[e][1]);
                acc[e][2] = __ffma2_rn(b2, *reinterpret_cast<const float2*>(te + 2 * Lp), acc[e][2]);
              }
            }
          }
        }
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          const float r = warp_sum(acc[e][0].x + acc[e][0].y), g = warp_sum(acc[e][1].x + acc[e][1].y),
                      b = warp_sum(acc[e][2].x + acc[e][2].y);
          if (lane == j) { outv[e][0] = r; outv[e][1] = g; outv[e][2] = b; }
        }
      }
      if (pv) {
#pragma unroll
        for (int e = 0; e < EC; ++e) {
          if (e < ec) {
            float* o = a.rgb_d + ((size_t)pi * a.n_envmaps + e0 + e) * 3;
            o[0] = tonemap(outv[e][0], a.linear2srgb);
            o[1] = tonemap(outv[e][1], a.linear2srgb);
            o[2] = tonemap(outv[e][2], a.linear2srgb);
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

// Microfacet.__call__ (brdf/microfacet/microfacet.py:30-72) as a standalone op: brdf[n, L, 3]
// for caller-supplied directions, every step in the reference's order (this is the drop-in for
// user code that calls the class directly; the renderer above never materialises [n, L, 3]).
__global__ void __launch_bounds__(256)
microfacet_brdf_kernel(const float* __restrict__ pts2l, const float* __restrict__ pts2c,
                       const float* __restrict__ normal, const float* __restrict__ albedo,
                       const float* __restrict__ rough, int n, int L, float default_rough,
                       int lambert_only, float f0, float* __restrict__ out) {
  const long long total = (long long)n * L;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / L);
    const f3 l = l2n(ld3(pts2l + idx * 3), 1e-6f);                 // :46
    const f3 v = l2n(ld3(pts2c + (size_t)i * 3), 1e-6f);           // :47
    const f3 nn = l2n(ld3(normal + (size_t)i * 3), 1e-6f);         // :48
    const f3 alb = albedo ? ld3(albedo + (size_t)i * 3) : mk3(1.f, 1.f, 1.f);   // :40-41
    const float r = rough ? rough[i] : default_rough;              // :42-44
    float micro = 0.f;
    if (!lambert_only) {
      const f3 h = l2n(l + v, 1e-6f);                              // :51-52
      const float om = 1.f - dot3(l, h);
      const float om2 = om * om;
      const float f = f0 + (1.f - f0) * (om2 * om2 * om);          // :106-111
      const float alpha = r * r;                                   // :54
      const float a2 = alpha * alpha;                              // alpha ** 2 (:88, :102)
      // _get_d (:93-104)
      const float cm = dot3(h, nn);
      const float chi_d = cm > 0.f ? 1.f : 0.f;
      const float cm2 = cm * cm;
      const float tan_m = divide_no_nan(1.f - cm2, cm2);
      const float t = a2 + tan_m;
      const float d = divide_no_nan(a2 * chi_d, NF_PI_F * (cm2 * cm2) * (t * t));
      // _get_g (:74-91)
      const float cos_v = dot3(nn, v);
      const float chi_g = divide_no_nan(dot3(h, v), cos_v) > 0.f ? 1.f : 0.f;
      const float cv2 = fminf(fmaxf(cos_v * cos_v, 0.f), 1.f);
      const float tan_v = fmaxf(divide_no_nan(1.f - cv2, cv2), 0.f);
      const float g = divide_no_nan(chi_g * 2.f, 1.f + sqrtf(1.f + a2 * tan_v));
      const float den = 4.f * fabsf(dot3(l, nn)) * fabsf(cos_v);  // :58-60
      micro = divide_no_nan(f * g * d, den);                       // :61
    }
    out[idx * 3 + 0] = micro + alb.x / NF_PI_F;                    // :64-71
    out[idx * 3 + 1] = micro + alb.y / NF_PI_F;
    out[idx * 3 + 2] = micro + alb.z / NF_PI_F;
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
  // env-maps per pass: as few passes over the lvis rows as E_CHUNK allows, split evenly
  // (E = 9 -> 2 passes of 5, run by the EC = 6 instantiation)
  const int n_pass = (a->n_envmaps + E_CHUNK - 1) / E_CHUNK;
  int ec = (a->n_envmaps + n_pass - 1) / n_pass;
  if (ec == 5) ec = 6;
  if (ec == 7) ec = 8;
  const size_t lp = ((size_t)a->n_lights + 1) & ~(size_t)1;
  size_t sm = sizeof(float) * lp * (4 + 3 * ec);
  // many lights AND many env-maps: fewer maps per pass until the texel arrays fit in shared memory
  static const int kLower[9] = {0, 1, 1, 2, 3, 4, 4, 6, 6};
  while (sm > ctx->smem_optin && ec > 1) {
    ec = kLower[ec];
    sm = sizeof(float) * lp * (4 + 3 * ec);
  }
  NF_CHECK_ARG(ctx, sm <= ctx->smem_optin, "n_lights too large for shared memory");
  // resident grid (blocks per SM from the occupancy calculator): every warp then walks ~8 batches
  // of 32 points at the full 800 x 800 size, keeping the tail imbalance small
  const int blocks_needed = (a->n + WARPS * 32 - 1) / (WARPS * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define NF_LAUNCH_INTEGRATE_K(EC, KIND)                                                           \
  do {                                                                                            \
    NF_CUDA(ctx, cudaFuncSetAttribute(integrate_kernel<EC, KIND>,                                 \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));    \
    int per_sm = 1;                                                                               \
    NF_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(                                   \
                     &per_sm, integrate_kernel<EC, KIND>, WARPS * 32, sm));                       \
    int grid = ctx->sm_count * (per_sm > 0 ? per_sm : 1);                                         \
    if (grid > blocks_needed) grid = blocks_needed;                                               \
    integrate_kernel<EC, KIND><<<grid, WARPS * 32, sm, st>>>(*a);                                 \
  } while (0)
#define NF_LAUNCH_INTEGRATE(EC)                                                                   \
  do {                                                                                            \
    if (a->brdf_kind == 0) NF_LAUNCH_INTEGRATE_K(EC, 0);                                          \
    else NF_LAUNCH_INTEGRATE_K(EC, 1);                                                            \
  } while (0)
  switch (ec) {
    case 1: NF_LAUNCH_INTEGRATE(1); break;
    case 2: NF_LAUNCH_INTEGRATE(2); break;
    case 3: NF_LAUNCH_INTEGRATE(3); break;
    case 4: NF_LAUNCH_INTEGRATE(4); break;
    case 6: NF_LAUNCH_INTEGRATE(6); break;
    default: NF_LAUNCH_INTEGRATE(8); break;
  }
#undef NF_LAUNCH_INTEGRATE
#undef NF_LAUNCH_INTEGRATE_K
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

int nf_microfacet_brdf_fwd(nf_ctx* ctx, const float* pts2l_d, const float* pts2c_d,
                           const float* normal_d, const float* albedo_d, const float* rough_d,
                           int n, int n_lights, float default_rough, int lambert_only, float f0,
                           float* brdf_d, void* stream) {
  NF_CHECK_ARG(ctx, n >= 0 && n_lights > 0, "bad n / n_lights");
  if (n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, pts2l_d && pts2c_d && normal_d && brdf_d, "null buffer");
  const long long total = (long long)n * n_lights;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  microfacet_brdf_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      pts2l_d, pts2c_d, normal_d, albedo_d, rough_d, n, n_lights, default_rough, lambert_only, f0,
      brdf_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

}  // extern "C"
