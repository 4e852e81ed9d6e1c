// Stage A support kernels: ray generation, stratified depths, volume-rendering
// weights / reductions, inverse-CDF resampling, light-visibility ray set-up.
#include <math.h>

#include "nf_common.cuh"

namespace {

// ------------------------------------------------------------------ gen_rays
// datasets/nerf.py:176-193 in fp64 without FMA contraction so the fp32 result is
// bit-identical to the NumPy path: d_local = ((x-.5W)/fl, -(y-.5H)/fl, -1),
// d_world[j] = (d0*R[j][0] + d1*R[j][1]) + d2*R[j][2].
__global__ void gen_rays_kernel(int h, int w, double fl, double r00, double r01, double r02,
                                double r10, double r11, double r12, double r20, double r21,
                                double r22, float ox, float oy, float oz, int normalize,
                                float* __restrict__ rayo, float* __restrict__ rayd) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= h * w) return;
  int y = n / w, x = n % w;   // ray n = y * W + x (nerf.py:109-110)
  double d0 = __ddiv_rn(__dsub_rn((double)x, __dmul_rn(.5, (double)w)), fl);
  double d1 = -__ddiv_rn(__dsub_rn((double)y, __dmul_rn(.5, (double)h)), fl);
  double d2 = -1.0;
  double wx = __dadd_rn(__dadd_rn(__dmul_rn(d0, r00), __dmul_rn(d1, r01)), __dmul_rn(d2, r02));
  double wy = __dadd_rn(__dadd_rn(__dmul_rn(d0, r10), __dmul_rn(d1, r11)), __dmul_rn(d2, r12));
  double wz = __dadd_rn(__dadd_rn(__dmul_rn(d0, r20), __dmul_rn(d1, r21)), __dmul_rn(d2, r22));
  f3 d = mk3((float)wx, (float)wy, (float)wz);
  if (normalize) d = l2n(d, 1e-12f);   // geometry_from_nerf.py:100
  rayo[n * 3 + 0] = ox; rayo[n * 3 + 1] = oy; rayo[n * 3 + 2] = oz;
  rayd[n * 3 + 0] = d.x; rayd[n * 3 + 1] = d.y; rayd[n * 3 + 2] = d.z;
}

// --------------------------------------------------------------------- gen_z
__device__ __forceinline__ float z_at(float near, float far, float step, int i, int lin_in_disp) {
  float t = 0.f + (float)i * step;                                 // tf.linspace(0., 1., S)
  if (lin_in_disp) return 1.f / (1.f / near * (1.f - t) + 1.f / far * t);
  return near * (1.f - t) + far * t;                               // nerf.py:122-126
}

__global__ void gen_z_kernel(float near, float far, int S, long long total, int lin_in_disp,
                             const float* __restrict__ u, float* __restrict__ z) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  int s = (int)(g % S);
  float step = S > 1 ? (1.f - 0.f) / (float)(S - 1) : 0.f;
  float zc = z_at(near, far, step, s, lin_in_disp);
  if (u) {                                                          // nerf.py:129-135
    float zl = s > 0 ? z_at(near, far, step, s - 1, lin_in_disp) : zc;
    float zr = s < S - 1 ? z_at(near, far, step, s + 1, lin_in_disp) : zc;
    float lower = s > 0 ? .5f * (zc + zl) : zc;
    float upper = s < S - 1 ? .5f * (zr + zc) : zc;
    zc = lower + (upper - lower) * u[g];
  }
  z[g] = zc;
}

// ----------------------------------------------------------------- composite
// One warp per ray; exclusive cumprod of (1 - alpha + 1e-6) by warp scan with a
// running carry (nerf.py:188-211, util/math.py:67-68).
__global__ void __launch_bounds__(256) composite_kernel(
    const float* __restrict__ sigma, const float* __restrict__ z,
    const float* __restrict__ rayo, const float* __restrict__ rayd,
    const float* __restrict__ normal, int n_rays, int S, float* __restrict__ weights,
    float* __restrict__ occu, float* __restrict__ depth, float* __restrict__ surf,
    float* __restrict__ exp_normal) {
  int ray = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (ray >= n_rays) return;
  f3 d = ld3(rayd + (size_t)ray * 3);
  float dn = sqrtf(dot3(d, d));                                    // tf.linalg.norm(rayd)
  const float* sg = sigma + (size_t)ray * S;
  const float* zz = z + (size_t)ray * S;
  float carry = 1.f, s_occu = 0.f, s_depth = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    int s = s0 + lane;
    float alpha = 0.f, zs = 0.f, tt = 1.f;
    if (s < S) {
      zs = zz[s];
      float dist = (s + 1 < S) ? (zz[s + 1] - zs) : 1e10f;
      dist *= dn;
      alpha = 1.0f - expf(-fmaxf(sg[s], 0.f) * dist);
      tt = 1.f - alpha + 1e-6f;
    }
    // inclusive product scan of tt
    float inc = tt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc *= v;
    }
    float exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 1.f;
    float T = carry * exc;
    float w = alpha * T;
    carry *= __shfl_sync(0xffffffffu, inc, 31);
    if (s < S) {
      if (weights) weights[(size_t)ray * S + s] = w;
      s_occu += w;
      s_depth += w * zs;
      if (normal) {
        const float* nn = normal + ((size_t)ray * S + s) * 3;
        nx += w * nn[0]; ny += w * nn[1]; nz += w * nn[2];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_occu += __shfl_xor_sync(0xffffffffu, s_occu, o);
    s_depth += __shfl_xor_sync(0xffffffffu, s_depth, o);
    nx += __shfl_xor_sync(0xffffffffu, nx, o);
    ny += __shfl_xor_sync(0xffffffffu, ny, o);
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
  }
  if (lane == 0) {
    if (occu) occu[ray] = s_occu;
    if (depth) depth[ray] = s_depth;
    if (surf) {                                                    // gfn.py:134
      f3 o = ld3(rayo + (size_t)ray * 3);
      surf[ray * 3 + 0] = o.x + d.x * s_depth;
      surf[ray * 3 + 1] = o.y + d.y * s_depth;
      surf[ray * 3 + 2] = o.z + d.z * s_depth;
    }
    if (exp_normal) { exp_normal[ray * 3] = nx; exp_normal[ray * 3 + 1] = ny; exp_normal[ray * 3 + 2] = nz; }
  }
}

// ---------------------------------------------------------------- gen_z_fine
// One warp per ray.  smem per warp: cdf[Sc-1], mid[Sc-1], zf[Sf].
__global__ void __launch_bounds__(128) gen_z_fine_kernel(
    const float* __restrict__ zc, const float* __restrict__ w, int n_rays, int Sc, int Sf,
    float* __restrict__ zall) {
  extern __shared__ float sm[];
  const int wpb = blockDim.x >> 5, wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * wpb + wi;
  const int nb = Sc - 1;                  // cdf / mid length (math.py:77-78, nerf.py:139)
  float* cdf = sm + (size_t)wi * (2 * nb + Sf);
  float* mid = cdf + nb;
  float* zf = mid + nb;
  if (ray >= n_rays) return;
  const float* zr = zc + (size_t)ray * Sc;
  const float* wr = w + (size_t)ray * Sc;
  // denom = sum(w[1:-1]) + eps  (math.py:72-73)
  float ssum = 0.f;
  for (int i = 1 + lane; i < Sc - 1; i += 32) ssum += wr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
  const float denom = ssum + 1e-5f;
  // cdf[0] = 0, cdf[k] = sum_{i<=k} pdf[i], pdf[i] = w[i+1-... ] : pdf index j <-> w[j+1]
  float carry = 0.f;
  for (int k0 = 0; k0 < nb; k0 += 32) {
    int k = k0 + lane;                     // cdf index k >= 1 uses pdf[k-1] = w[k] / denom
    float pv = (k >= 1 && k < nb) ? wr[k] / denom : 0.f;
    float inc = pv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (k < nb) {
      cdf[k] = carry + inc;
      mid[k] = .5f * (zr[k + 1] + zr[k]);
    }
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  __syncwarp();
  const float ustep = Sf > 1 ? 1.f / (float)(Sf - 1) : 0.f;
  for (int j = lane; j < Sf; j += 32) {
    float u = 0.f + (float)j * ustep;      // tf.linspace(0., 1., n_samples)
    // searchsorted side='right': first i with cdf[i] > u
    int lo = 0, hi = nb;
    while (lo < hi) {
      int m = (lo + hi) >> 1;
      if (cdf[m] > u) hi = m; else lo = m + 1;
    }
    int ind = lo;
    int below = max(0, ind - 1), above = min(ind, nb - 1);
    float cb = cdf[below], ca = cdf[above];
    float dd = ca - cb;
    if (dd < 1e-5f) dd = 1.f;
    float t = (u - cb) / dd;
    zf[j] = mid[below] + t * (mid[above] - mid[below]);
  }
  __syncwarp();
  // merge the two sorted lists (tf.sort of the concat, nerf.py:145-146)
  float* out = zall + (size_t)ray * (Sc + Sf);
  for (int i = lane; i < Sc; i += 32) {
    float v = zr[i];
    int lo = 0, hi = Sf;                    // # of zf strictly less than v
    while (lo < hi) { int m = (lo + hi) >> 1; if (zf[m] < v) lo = m + 1; else hi = m; }
    out[i + lo] = v;
  }
  for (int j = lane; j < Sf; j += 32) {
    float v = zf[j];
    int lo = 0, hi = Sc;                    // # of zc less than or equal to v
    while (lo < hi) { int m = (lo + hi) >> 1; if (zr[m] <= v) lo = m + 1; else hi = m; }
    out[j + lo] = v;
  }
}

__global__ void lvis_rays_kernel(const float* __restrict__ surf, const float* __restrict__ normal,
                                 long long total, int L, const float* __restrict__ lxyz,
                                 float* __restrict__ rayo, float* __restrict__ rayd,
                                 uint8_t* __restrict__ fl) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  long long i = g / L;
  int l = (int)(g % L);
  f3 s = ld3(surf + i * 3);
  f3 d = l2n(ld3(lxyz + l * 3) - s, 1e-12f);                       // gfn.py:197-198
  float c = dot3(d, ld3(normal + i * 3));                          // gfn.py:205-206
  rayo[g * 3] = s.x; rayo[g * 3 + 1] = s.y; rayo[g * 3 + 2] = s.z;
  rayd[g * 3] = d.x; rayd[g * 3 + 1] = d.y; rayd[g * 3 + 2] = d.z;
  fl[g] = c > 0.f ? 1 : 0;
}

// Input rows of the light-visibility network for training-size batches, materialised in one
// launch: row (i, l) = [embed(xyz_scale * xyz_i) | embed(l2n(lxyz_l - xyz_dir_i)) | 0-padding]
// (shape.py:128-135, 213-233; embedder.py:46-47).  One thread per (row, octave-or-raw) triple.
__global__ void __launch_bounds__(256)
lvis_inputs_kernel(const float* __restrict__ xyz, const float* __restrict__ xyz_dir,
                   const float* __restrict__ lxyz, long long rows, int L, float xyz_scale, int fa,
                   int fb, int ld, float* __restrict__ out) {
  const int per_row = (1 + fa) + (1 + fb);          // items: raw + octaves of each encoding
  const long long total = rows * per_row;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / per_row;
    int item = (int)(idx % per_row);
    const long long i = r / L;
    const int l = (int)(r % L);
    float* o = out + r * ld;
    f3 v;
    int base, k;
    if (item < 1 + fa) {                             // position encoding
      v = ld3(xyz + i * 3) * xyz_scale;
      base = 0; k = item;
    } else {                                         // light-direction encoding
      v = l2n(ld3(lxyz + l * 3) - ld3(xyz_dir + i * 3), 1e-6f);
      base = 3 * (1 + 2 * fa); k = item - (1 + fa);
      if (k == 0)                                    // zero padding up to the row stride
        for (int c = base + 3 * (1 + 2 * fb); c < ld; ++c) o[c] = 0.f;
    }
    if (k == 0) { o[base] = v.x; o[base + 1] = v.y; o[base + 2] = v.z; continue; }
    const float fr = (float)(1 << (k - 1));
    float* q = o + base + 3 + 6 * (k - 1);
    q[0] = sinf(v.x * fr); q[1] = sinf(v.y * fr); q[2] = sinf(v.z * fr);
    q[3] = cosf(v.x * fr); q[4] = cosf(v.y * fr); q[5] = cosf(v.z * fr);
  }
}

}  // namespace

extern "C" {

int nf_lvis_inputs_fwd(nf_ctx* ctx, const float* xyz_d, const float* xyz_dir_d, int n,
                       const float* lxyz_d, int n_lights, float xyz_scale, int n_freqs_xyz,
                       int n_freqs_ldir, int ld, float* out_d, void* stream) {
  NF_CHECK_ARG(ctx, xyz_d && xyz_dir_d && lxyz_d && out_d, "null argument");
  NF_CHECK_ARG(ctx, n >= 0 && n_lights > 0 && n_freqs_xyz >= 0 && n_freqs_ldir >= 0, "bad sizes");
  NF_CHECK_ARG(ctx, ld >= 3 * (1 + 2 * n_freqs_xyz) + 3 * (1 + 2 * n_freqs_ldir), "row stride too small");
  const long long rows = (long long)n * n_lights;
  if (rows == 0) return NF_OK;
  const long long total = rows * (2 + n_freqs_xyz + n_freqs_ldir);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 32;
  if (blocks > cap) blocks = cap;
  lvis_inputs_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      xyz_d, xyz_dir_d, lxyz_d, rows, n_lights, xyz_scale, n_freqs_xyz, n_freqs_ldir, ld, out_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_gen_rays(nf_ctx* ctx, const double* c2w, double cam_angle_x, int h, int w,
                int normalize, float* rayo_d, float* rayd_d, void* stream) {
  NF_CHECK_ARG(ctx, c2w && rayo_d && rayd_d, "null argument");
  NF_CHECK_ARG(ctx, h > 0 && w > 0 && (long long)h * w < 2147483647LL, "bad image size");
  double fl = .5 * w / tan(.5 * cam_angle_x);                      // nerf.py:186 (host libm, like NumPy)
  int n = h * w;
  gen_rays_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      h, w, fl, c2w[0], c2w[1], c2w[2], c2w[4], c2w[5], c2w[6], c2w[8], c2w[9], c2w[10],
      (float)c2w[3], (float)c2w[7], (float)c2w[11], normalize, rayo_d, rayd_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_gen_z(nf_ctx* ctx, float near, float far, int n_samples, int n_rays, int lin_in_disp,
             const float* perturb_u_d, float* z_d, void* stream) {
  NF_CHECK_ARG(ctx, z_d && n_samples > 0 && n_rays >= 0, "bad argument");
  long long total = (long long)n_rays * n_samples;
  if (total == 0) return NF_OK;
  gen_z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      near, far, n_samples, total, lin_in_disp, perturb_u_d, z_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_composite(nf_ctx* ctx, const float* sigma_d, const float* z_d, const float* rayo_d,
                 const float* rayd_d, const float* normal_d, int n_rays, int n_samples,
                 float* weights_d, float* occu_d, float* depth_d, float* surf_d,
                 float* exp_normal_d, void* stream) {
  NF_CHECK_ARG(ctx, sigma_d && z_d && rayd_d && n_samples > 0, "bad argument");
  NF_CHECK_ARG(ctx, !surf_d || rayo_d, "surf needs rayo");
  NF_CHECK_ARG(ctx, !exp_normal_d || normal_d, "exp_normal needs normal");
  if (n_rays == 0) return NF_OK;
  composite_kernel<<<(n_rays + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
      sigma_d, z_d, rayo_d, rayd_d, normal_d, n_rays, n_samples, weights_d, occu_d, depth_d,
      surf_d, exp_normal_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_gen_z_fine(nf_ctx* ctx, const float* z_coarse_d, const float* weights_d, int n_rays,
                  int n_coarse, int n_fine, float* z_all_d, void* stream) {
  NF_CHECK_ARG(ctx, z_coarse_d && weights_d && z_all_d, "null argument");
  NF_CHECK_ARG(ctx, n_coarse >= 3 && n_fine >= 1, "need n_coarse >= 3, n_fine >= 1");
  if (n_rays == 0) return NF_OK;
  const int wpb = 4;
  size_t sm = sizeof(float) * wpb * (size_t)(2 * (n_coarse - 1) + n_fine);
  NF_CHECK_ARG(ctx, sm <= 48 * 1024, "too many samples per ray");
  gen_z_fine_kernel<<<(n_rays + wpb - 1) / wpb, wpb * 32, sm, (cudaStream_t)stream>>>(
      z_coarse_d, weights_d, n_rays, n_coarse, n_fine, z_all_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

int nf_lvis_rays(nf_ctx* ctx, const float* surf_d, const float* normal_d, int n_pts,
                 const float* lxyz_d, int n_lights, float* rayo_d, float* rayd_d,
                 uint8_t* front_lit_d, void* stream) {
  NF_CHECK_ARG(ctx, surf_d && normal_d && lxyz_d && rayo_d && rayd_d && front_lit_d, "null argument");
  long long total = (long long)n_pts * n_lights;
  if (total == 0) return NF_OK;
  lvis_rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      surf_d, normal_d, total, n_lights, lxyz_d, rayo_d, rayd_d, front_lit_d);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

}  // extern "C"
