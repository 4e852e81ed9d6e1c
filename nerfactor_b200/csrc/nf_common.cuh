// Shared declarations for the nerfactor_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/nerfactor_b200.h"

struct nf_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t smem_optin = 0;
  std::string last_error;
};

// Host-side handle of one packed network (see nf_mlp_create in nf_api.cu).
struct nf_mlp {
  nf_mlp_desc d;              // copy of the descriptor (W / b pointers cleared)
  int in_pad = 0;             // in_dim rounded up to a multiple of 16
  std::vector<uint8_t> blob;  // host image of the device buffer
  // byte offsets into the blob / device buffer
  std::vector<size_t> off_w32;  // [depth+1] fp32 W[l]  ([K_l][out_l], Keras layout)
  std::vector<size_t> off_b32;  // [depth+1] fp32 b[l]
  // tensor-core images (built for the kinds that have a tcgen05 kernel)
  size_t off_tc_f16 = 0, off_tc_bf16 = 0, tc_bytes = 0;  // K-major operand-B images
  size_t off_tc_aux = 0, tc_aux_bytes = 0;               // fp32 side blocks
  size_t off_tc2_f16 = 0, off_tc2_bf16 = 0;              // CTA-pair (cta_group::2) images
  size_t off_tcb_f16 = 0, off_tcb_bf16 = 0;              // sigma nets: backward (d sigma / d x) images
  size_t off_rgb_f16 = 0, off_rgb_bf16 = 0, off_rgb_aux = 0;   // NeRF colour branch (nf_mlp_attach_rgb)
  std::vector<size_t> off_wt;   // sigma nets: transposed trunk weights for d sigma / d xyz
  std::vector<int> ldk;
  void* dev = nullptr;        // caller-owned device buffer (after nf_mlp_upload)
};

int nf_set_error(nf_ctx* ctx, int code, const char* fmt, ...);

#define NF_CHECK_ARG(ctx, cond, msg)                                        \
  do {                                                                      \
    if (!(cond)) return nf_set_error((ctx), NF_ERR_INVALID_ARG, "%s: %s", __func__, (msg)); \
  } while (0)

#define NF_CUDA(ctx, call)                                                  \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess)                                                 \
      return nf_set_error((ctx), NF_ERR_CUDA, "%s: %s (%s:%d)", __func__,   \
                          cudaGetErrorString(e__), __FILE__, __LINE__);     \
  } while (0)

#define NF_LAUNCH_CHECK(ctx) NF_CUDA(ctx, cudaGetLastError())

// ------------------------------------------------------------------ device math
#define NF_PI_F 3.14159265358979323846f

struct f3 {
  float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }

// tf.linalg.l2_normalize(x, eps): x * rsqrt(max(sum x^2, eps))   (util/math.py:63-64)
__device__ __forceinline__ f3 l2n(f3 v, float eps) {
  float s = 1.0f / sqrtf(fmaxf(dot3(v, v), eps));
  return v * s;
}
__device__ __forceinline__ float divide_no_nan(float a, float b) {
  return b == 0.f ? 0.f : a / b;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf_(float x) {
  // log(1 + e^x), overflow-safe
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float apply_act(int act, float x) {
  switch (act) {
    case NF_ACT_RELU: return fmaxf(x, 0.f);
    case NF_ACT_SIGMOID: return sigmoidf_(x);
    case NF_ACT_SOFTPLUS: return softplusf_(x);
    default: return x;
  }
}
// nerfactor/util/img.py:140-163 on an already [0,1]-clipped value
__device__ __forceinline__ float linear2srgb_dev(float x) {
  float lin = x * 12.92f;
  float nonlin = 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
  return x <= 0.0031308f ? lin : nonlin;
}

// nerfactor/util/geom.py:119-149 -- rows (t, b, n) of world2local
__device__ __forceinline__ void world2local_dev(f3 normal, f3& t, f3& b, f3& n) {
  n = l2n(normal, 1e-6f);
  f3 z = mk3(0.f + 1e-6f, 0.f + 1e-6f, 1.f + 1e-6f);
  t = l2n(cross3(n, z), 1e-6f);
  b = l2n(cross3(n, t), 1e-6f);
}

// Rodrigues rotation, same operation order as geom.py:167-181 (rot_vec)
__device__ __forceinline__ f3 rot_vec_dev(f3 v, f3 axis, float angle) {
  float s, c;
  sincosf(angle, &s, &c);
  float va = dot3(v, axis);
  f3 cr = cross3(axis, v);
  return mk3(v.x * c + axis.x * va * (1.f - c) + cr.x * s,
             v.y * c + axis.y * va * (1.f - c) + cr.y * s,
             v.z * c + axis.z * va * (1.f - c) + cr.z * s);
}

// nerfactor/util/geom.py:152-192 -- returns (phi_d, theta_h, theta_d)
__device__ __forceinline__ f3 dir2rusink_dev(f3 a, f3 b) {
  a = l2n(a, 1e-6f);
  b = l2n(b, 1e-6f);
  f3 h = l2n((a + b) * 0.5f, 1e-6f);
  float theta_h = acosf(fminf(fmaxf(h.z, -1.f), 1.f));
  float phi_h = atan2f(h.y, h.x);
  f3 tmp = rot_vec_dev(b, mk3(0.f, 0.f, 1.f), -phi_h);
  f3 diff = rot_vec_dev(tmp, mk3(0.f, 1.f, 0.f), -theta_h);
  float theta_d = acosf(fminf(fmaxf(diff.z, -1.f), 1.f));
  float at = atan2f(diff.y, diff.x);
  float phi_d = at - floorf(at / NF_PI_F) * NF_PI_F;  // tf.math.floormod(., pi)
  return mk3(phi_d, theta_h, theta_d);
}

static inline int nf_round_up(int x, int m) { return (x + m - 1) / m * m; }
