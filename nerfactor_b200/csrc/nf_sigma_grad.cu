// sigma and its input gradient: normal = -l2_normalize(d relu(sigma_raw) / d xyz)
// (the GradientTape.batch_jacobian block of nerfactor/geometry_from_nerf.py:285-300).
//
// FP32 CUDA-core kernel, 64 samples per CTA.  Forward = the fused skip-MLP of
// nf_mlp_simt.cu with the ReLU pattern of every layer kept as a bit mask in shared
// memory; backward = the same tiled GEMM against the pre-transposed weights
// (g_in = (g_out * relu') W^T), the skip layer's input block and layer 0 accumulate
// the gradient w.r.t. the positional encoding, and the chain rule through
// [x, sin(2^f x), cos(2^f x)] (nerfactor/networks/embedder.py:46-47) finishes in registers.
#include "nf_common.cuh"

namespace {

constexpr int TILE_M = 64;
constexpr int KC = 16;
constexpr int X_STRIDE = 100;
constexpr int NTHREADS = 256;
constexpr int WIDTH = 256;
constexpr int H_STRIDE = WIDTH + 4;
constexpr int GE_STRIDE = 68;
constexpr int TXN = WIDTH / 8;          // 32
constexpr int TM = TILE_M / (NTHREADS / TXN);   // 8

struct GradParams {
  const uint8_t* blob;
  size_t off_w[9], off_b[9], off_wt[8];
  int ldk[8];
  int in_dim, in_pad, depth, skip_at, n_freqs;
  long long n_rows;
  int per;
  const float* rayo;
  const float* rayd;
  const float* z;
  float bbox[6];
  int use_bbox;
  float* sigma;
  float* normal;
};

__device__ __forceinline__ void cp16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[TM][8] += A[64 x kreal] (smem, stride as) * W[kreal x (cols col0..col0+255)] (global,
// row stride ldw); columns >= ncols_valid read as zero.
__device__ __forceinline__ void gemm_pass(const float* A, int as, const float* Wg, int ldw,
                                          int kreal, int col0, int ncols_valid, float* Wc,
                                          float (&acc)[TM][8], int tid, int tx, int ty) {
  const int nchunk = (kreal + KC - 1) / KC;
  auto load_chunk = [&](int c, int buf) {
    for (int i = tid; i < KC * WIDTH / 4; i += NTHREADS) {
      int kk = i / (WIDTH / 4), c4 = i % (WIDTH / 4);
      int krow = c * KC + kk;
      bool valid = krow < kreal && (c4 * 4) < ncols_valid;
      const float* src = Wg + (size_t)(valid ? krow : 0) * ldw + col0 + (valid ? c4 * 4 : 0);
      cp16(Wc + buf * KC * WIDTH + kk * WIDTH + c4 * 4, src, valid);
    }
    cp_commit();
  };
  load_chunk(0, 0);
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) { load_chunk(c + 1, (c + 1) & 1); cp_wait<1>(); }
    else cp_wait<0>();
    __syncthreads();
    const float* W = Wc + (c & 1) * KC * WIDTH;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i)
        a[i] = *reinterpret_cast<const float4*>(A + (ty * TM + i) * as + c * KC + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 w0 = *reinterpret_cast<const float4*>(W + (k4 + kk) * WIDTH + tx * 4);
        float4 w1 = *reinterpret_cast<const float4*>(W + (k4 + kk) * WIDTH + WIDTH / 2 + tx * 4);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
          acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
          acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NTHREADS) sigma_grad_kernel(const GradParams p) {
  extern __shared__ __align__(16) float smem[];
  float* X0 = smem;                                   // [64][100] embedding (kept)
  float* H0 = X0 + TILE_M * X_STRIDE;                 // [64][260]
  float* H1 = H0 + TILE_M * H_STRIDE;
  float* Wc = H1 + TILE_M * H_STRIDE;                 // [2][16][256]
  float* GE = Wc + 2 * KC * WIDTH;                    // [64][68] grad w.r.t. embedding
  unsigned* MK = reinterpret_cast<unsigned*>(GE + TILE_M * GE_STRIDE);   // [8][64][8] ReLU masks
  float* s_pos = reinterpret_cast<float*>(MK + 8 * TILE_M * 8);          // [64] raw > 0
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
  const long long row0 = (long long)blockIdx.x * TILE_M;

  for (int i = tid; i < 8 * TILE_M * 8; i += NTHREADS) MK[i] = 0u;
  // ---------------------------------------------------------------- prologue
  {
    const int r = tid >> 2, sub = tid & 3;
    const long long g = row0 + r;
    const bool live = g < p.n_rows;
    float* xr = X0 + r * X_STRIDE;
    if (sub == 3) for (int k = p.in_dim; k < p.in_pad; ++k) xr[k] = 0.f;
    long long ray = live ? g / p.per : 0;
    float zz = live ? p.z[g] : 0.f;
    if (sub < 3) {
      float pc = __fadd_rn(p.rayo[ray * 3 + sub], __fmul_rn(p.rayd[ray * 3 + sub], zz));  // gfn.py:264
      xr[sub] = pc;
      float f = 1.f;
      for (int k = 0; k < p.n_freqs; ++k) {
        float s, co;
        sincosf(pc * f, &s, &co);
        xr[3 + 6 * k + sub] = s;
        xr[3 + 6 * k + 3 + sub] = co;
        f *= 2.f;
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- forward
  float* Hin = nullptr;
  float* Hout = H0;
  for (int l = 0; l < p.depth; ++l) {
    const float* Wg = reinterpret_cast<const float*>(p.blob + p.off_w[l]);
    const float* bg = reinterpret_cast<const float*>(p.blob + p.off_b[l]);
    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    if (l == 0) gemm_pass(X0, X_STRIDE, Wg, WIDTH, p.in_dim, 0, WIDTH, Wc, acc, tid, tx, ty);
    else {
      gemm_pass(Hin, H_STRIDE, Wg, WIDTH, WIDTH, 0, WIDTH, Wc, acc, tid, tx, ty);
      if (l == p.skip_at + 1)
        gemm_pass(X0, X_STRIDE, Wg + (size_t)WIDTH * WIDTH, WIDTH, p.in_dim, 0, WIDTH, Wc, acc, tid, tx, ty);
    }
    float4 b0 = *reinterpret_cast<const float4*>(bg + tx * 4);
    float4 b1 = *reinterpret_cast<const float4*>(bg + WIDTH / 2 + tx * 4);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      float v[8] = {acc[i][0] + b0.x, acc[i][1] + b0.y, acc[i][2] + b0.z, acc[i][3] + b0.w,
                    acc[i][4] + b1.x, acc[i][5] + b1.y, acc[i][6] + b1.z, acc[i][7] + b1.w};
      unsigned m0 = 0u, m1 = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (v[j] > 0.f) m0 |= 1u << j;
        if (v[4 + j] > 0.f) m1 |= 1u << j;
        v[j] = fmaxf(v[j], 0.f);
        v[4 + j] = fmaxf(v[4 + j], 0.f);
      }
      const int row = ty * TM + i;
      const int c0 = tx * 4, c1 = WIDTH / 2 + tx * 4;
      atomicOr(&MK[(l * TILE_M + row) * 8 + (c0 >> 5)], m0 << (c0 & 31));
      atomicOr(&MK[(l * TILE_M + row) * 8 + (c1 >> 5)], m1 << (c1 & 31));
      *reinterpret_cast<float4*>(Hout + row * H_STRIDE + c0) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(Hout + row * H_STRIDE + c1) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    Hin = Hout;
    Hout = (Hout == H0) ? H1 : H0;
  }

  // ---------------------------------------------------------------- head
  const float* Wo = reinterpret_cast<const float*>(p.blob + p.off_w[p.depth]);   // [256][1]
  const float* bo = reinterpret_cast<const float*>(p.blob + p.off_b[p.depth]);
  {
    const int r = tid >> 2, q = tid & 3;
    float o = 0.f;
    for (int c = q; c < WIDTH; c += 4) o = fmaf(Hin[r * H_STRIDE + c], Wo[c], o);
    o += __shfl_xor_sync(0xffffffffu, o, 1);
    o += __shfl_xor_sync(0xffffffffu, o, 2);
    const long long g = row0 + r;
    if (q == 0) {
      float raw = o + bo[0];
      s_pos[r] = raw > 0.f ? 1.f : 0.f;
      if (g < p.n_rows) {
        float v = fmaxf(raw, 0.f);                                  // tf.nn.relu, gfn.py:291-292
        if (p.use_bbox) {                                           // gfn.py:275-277, 303-305
          const long long ray = g / p.per;
          const float zz = p.z[g];
          const float px = __fadd_rn(p.rayo[ray * 3 + 0], __fmul_rn(p.rayd[ray * 3 + 0], zz));
          const float py = __fadd_rn(p.rayo[ray * 3 + 1], __fmul_rn(p.rayd[ray * 3 + 1], zz));
          const float pz = __fadd_rn(p.rayo[ray * 3 + 2], __fmul_rn(p.rayd[ray * 3 + 2], zz));
          const bool in = px >= p.bbox[0] && px <= p.bbox[1] && py >= p.bbox[2] &&
                          py <= p.bbox[3] && pz >= p.bbox[4] && pz <= p.bbox[5];
          if (!in) v = 0.f;
        }
        p.sigma[g] = v;
      }
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- backward
  // G = d raw / d h_depth = w_out, into the buffer that is free (Hout)
  float* Gcur = Hout;
  float* Gnext = Hin;
  for (int i = tid; i < TILE_M * WIDTH; i += NTHREADS) {
    int r = i / WIDTH, c = i % WIDTH;
    Gcur[r * H_STRIDE + c] = Wo[c];
  }
  __syncthreads();
  for (int l = p.depth - 1; l >= 0; --l) {
    // g_pre = g_h * relu'(pre_l)
    for (int i = tid; i < TILE_M * WIDTH; i += NTHREADS) {
      int r = i / WIDTH, c = i % WIDTH;
      if (!((MK[(l * TILE_M + r) * 8 + (c >> 5)] >> (c & 31)) & 1u)) Gcur[r * H_STRIDE + c] = 0.f;
    }
    __syncthreads();
    const float* WT = reinterpret_cast<const float*>(p.blob + p.off_wt[l]);    // [256][ldk]
    const int ldk = p.ldk[l];
    float acc[TM][8];
    if (l > 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      gemm_pass(Gcur, H_STRIDE, WT, ldk, WIDTH, 0, WIDTH, Wc, acc, tid, tx, ty);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int row = ty * TM + i;
        *reinterpret_cast<float4*>(Gnext + row * H_STRIDE + tx * 4) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(Gnext + row * H_STRIDE + WIDTH / 2 + tx * 4) =
            make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
    }
    if (l == 0 || l == p.skip_at + 1) {
      // gradient w.r.t. the embedding columns: WT columns [col0, col0 + 64)
      const int col0 = l == 0 ? 0 : WIDTH;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      gemm_pass(Gcur, H_STRIDE, WT, ldk, WIDTH, col0, p.in_pad, Wc, acc, tid, tx, ty);
      if (tx * 4 < p.in_pad) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          float* ge = GE + (ty * TM + i) * GE_STRIDE + tx * 4;
          if (l == 0) { ge[0] += acc[i][0]; ge[1] += acc[i][1]; ge[2] += acc[i][2]; ge[3] += acc[i][3]; }
          else { ge[0] = acc[i][0]; ge[1] = acc[i][1]; ge[2] = acc[i][2]; ge[3] = acc[i][3]; }
        }
      }
    }
    __syncthreads();
    float* t = Gcur; Gcur = Gnext; Gnext = t;
  }

  // ------------------------------------------- chain rule through the embedding
  if (tid < TILE_M) {
    const int r = tid;
    const long long g = row0 + r;
    if (g < p.n_rows) {
      const float* e = X0 + r * X_STRIDE;
      const float* ge = GE + r * GE_STRIDE;
      float gr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = ge[c];
        float f = 1.f;
        for (int k = 0; k < p.n_freqs; ++k) {
          // d sin(f x) = f cos(f x) dx ; d cos(f x) = -f sin(f x) dx
          a += f * (e[3 + 6 * k + 3 + c] * ge[3 + 6 * k + c] - e[3 + 6 * k + c] * ge[3 + 6 * k + 3 + c]);
          f *= 2.f;
        }
        gr[c] = a * s_pos[r];
      }
      const float s = 1.0f / sqrtf(fmaxf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2], 1e-12f));
      p.normal[g * 3 + 0] = -gr[0] * s;          // -tf.linalg.l2_normalize, gfn.py:297
      p.normal[g * 3 + 1] = -gr[1] * s;
      p.normal[g * 3 + 2] = -gr[2] * s;
    }
  }
}

constexpr size_t GRAD_SMEM = sizeof(float) * (size_t)(TILE_M * X_STRIDE + 2 * TILE_M * H_STRIDE +
                                                     2 * KC * WIDTH + TILE_M * GE_STRIDE +
                                                     8 * TILE_M * 8 + TILE_M);

}  // namespace

// Appends the transposed trunk weights WT_l[256][ldk_l] (ldk_l = K_l rounded up to 4) used by
// the backward GEMMs.  Called from nf_mlp_create for NF_MLP_SIGMA networks of width 256.
int nf_sigma_grad_pack(nf_mlp* m) {
  const nf_mlp_desc& d = m->d;
  if (d.kind != NF_MLP_SIGMA || d.width != 256 || d.depth > 8) return NF_OK;
  m->off_wt.assign(d.depth, 0);
  m->ldk.assign(d.depth, 0);
  for (int l = 0; l < d.depth; ++l) {
    const int K = l == 0 ? d.in_dim : (l == d.skip_at + 1 ? d.width + d.in_dim : d.width);
    const int ldk = (K + 3) / 4 * 4;
    size_t base = (m->blob.size() + 255) / 256 * 256;
    m->blob.resize(base + (size_t)256 * ldk * sizeof(float), 0);
    float* wt = reinterpret_cast<float*>(m->blob.data() + base);
    for (int k = 0; k < K; ++k)
      for (int j = 0; j < 256; ++j) wt[(size_t)j * ldk + k] = d.W[l][(size_t)k * 256 + j];
    m->off_wt[l] = base;
    m->ldk[l] = ldk;
  }
  return NF_OK;
}

int nf_tc_sigma_grad_launch(nf_ctx* ctx, const nf_mlp* m, const float* rayo, const float* rayd,
                            const float* z, int n_rays, int S, const float* bbox_host,
                            float* sigma, float* normal, int precision, cudaStream_t st);

extern "C" int nf_sigma_normal_fwd(nf_ctx* ctx, const nf_mlp* m, const float* rayo_d,
                                   const float* rayd_d, const float* z_d, int n_rays,
                                   int n_samples, const float* bbox_host, float* sigma_d,
                                   float* normal_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, m, "null network");
  NF_CHECK_ARG(ctx, m->d.kind == NF_MLP_SIGMA && m->d.out_dim == 1, "network is not NF_MLP_SIGMA");
  NF_CHECK_ARG(ctx, n_rays >= 0 && n_samples > 0, "bad sizes");
  if (n_rays == 0) return NF_OK;
  NF_CHECK_ARG(ctx, rayo_d && rayd_d && z_d && sigma_d && normal_d, "null buffer");
  NF_CHECK_ARG(ctx, m->dev, "network not uploaded (call nf_mlp_upload first)");
  NF_CHECK_ARG(ctx, precision == NF_PREC_FP32 || precision == NF_PREC_F16 || precision == NF_PREC_BF16,
               "precision must be NF_PREC_FP32, NF_PREC_F16 or NF_PREC_BF16");
  if (precision != NF_PREC_FP32)
    return nf_tc_sigma_grad_launch(ctx, m, rayo_d, rayd_d, z_d, n_rays, n_samples, bbox_host,
                                   sigma_d, normal_d, precision, (cudaStream_t)stream);
  if (m->off_wt.empty())
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "nf_sigma_normal_fwd needs a width-256 sigma network (depth <= 8)");
  GradParams p;
  memset(&p, 0, sizeof(p));
  p.blob = (const uint8_t*)m->dev;
  for (int l = 0; l <= m->d.depth; ++l) { p.off_w[l] = m->off_w32[l]; p.off_b[l] = m->off_b32[l]; }
  for (int l = 0; l < m->d.depth; ++l) { p.off_wt[l] = m->off_wt[l]; p.ldk[l] = m->ldk[l]; }
  p.in_dim = m->d.in_dim; p.in_pad = m->in_pad; p.depth = m->d.depth; p.skip_at = m->d.skip_at;
  p.n_freqs = m->d.n_freqs_a;
  p.n_rows = (long long)n_rays * n_samples; p.per = n_samples;
  p.rayo = rayo_d; p.rayd = rayd_d; p.z = z_d; p.sigma = sigma_d; p.normal = normal_d;
  if (bbox_host) { memcpy(p.bbox, bbox_host, sizeof(p.bbox)); p.use_bbox = 1; }
  const long long blocks = (p.n_rows + TILE_M - 1) / TILE_M;
  NF_CHECK_ARG(ctx, blocks < 2147483647LL, "too many samples for one launch");
  NF_CHECK_ARG(ctx, GRAD_SMEM <= ctx->smem_optin, "shared memory budget exceeded");
  NF_CUDA(ctx, cudaFuncSetAttribute(sigma_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)GRAD_SMEM));
  sigma_grad_kernel<<<(unsigned)blocks, NTHREADS, GRAD_SMEM, (cudaStream_t)stream>>>(p);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
