// Training-side kernels (FP32 CUDA cores): Dense forward / backward on materialised
// activations and the AMSGrad-Adam update.
//
// A train step of the reference (nerfactor/trainvali.py:273-286) touches 1024 rays x 512
// lights = 0.5 M network rows, three orders of magnitude fewer than a test-time view, so the
// layer-by-layer form with activations in HBM is adequate here; every kernel is the same
// 64-row x WIDTH-column register-tiled FFMA micro-kernel as nf_mlp_simt.cu.
//   forward   Y  = act([X1 | X2] W + b)                         (mlp.py:39-50, Keras Dense)
//   data grad dX = (dY * act'(Y)) W^T          -> [dX1 | dX2]
//   weight grad dW += [X1 | X2]^T (dY * act'(Y)),  db += colsum(dY * act'(Y))
#include "nf_common.cuh"

namespace {

constexpr int TILE_M = 64;
constexpr int KC = 16;
constexpr int NTHREADS = 256;

__device__ __forceinline__ void cp16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float act_grad(int act, float y) {
  switch (act) {
    case NF_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case NF_ACT_SIGMOID: return y * (1.f - y);
    case NF_ACT_SOFTPLUS: return 1.f - expf(-y);      // d/dx log(1+e^x) = 1 - e^{-softplus(x)}
    default: return 1.f;
  }
}

// C[64 x WIDTH tile] = A[64 x K] (global rows row0.., row stride lda, zero beyond M / K)
//                      * B[K x ldb] columns [col0, col0 + WIDTH) (zero beyond ncols)
template <int WIDTH>
struct Tile {
  static constexpr int TXN = WIDTH / 8;
  static constexpr int TYN = NTHREADS / TXN;
  static constexpr int TM = TILE_M / TYN;
};

// Generic pass: A chunk and B chunk both staged through shared memory per 16-wide K step.
// a_trans = 0: A(i, k) = A[(row0 + i) * lda + k]      (rows of A are GEMM rows)
// a_trans = 1: A(i, k) = A[(k) * lda + row0 + i]      (GEMM rows are columns of A: wgrad)
template <int WIDTH>
__device__ __forceinline__ void gemm_gg(const float* A, long long lda, int a_trans, long long arows,
                                        long long acols, long long row0, const float* B, int ldb,
                                        long long k0, long long k1, int col0, int ncols, float* As,
                                        float* Bs, float (&acc)[Tile<WIDTH>::TM][8], int tid) {
  constexpr int TM = Tile<WIDTH>::TM;
  const int tx = tid % Tile<WIDTH>::TXN, ty = tid / Tile<WIDTH>::TXN;
  // As: [2][TILE_M][KC + 4], Bs: [2][KC][WIDTH]
  constexpr int AS = KC + 4;
  auto load = [&](long long kb, int buf) {
    // A chunk: 64 x 16
    for (int i = tid; i < TILE_M * KC; i += NTHREADS) {
      int r, kk;
      if (a_trans) { kk = i / TILE_M; r = i % TILE_M; }   // consecutive threads -> consecutive rows
      else { r = i / KC; kk = i % KC; }
      long long gi = row0 + r, gk = kb + kk;
      float v = 0.f;
      if (a_trans) { if (gk < k1 && gi < acols) v = A[gk * lda + gi]; }
      else { if (gi < arows && gk < k1) v = A[gi * lda + gk]; }
      As[(buf * TILE_M + r) * AS + kk] = v;
    }
    for (int i = tid; i < KC * WIDTH / 4; i += NTHREADS) {
      int kk = i / (WIDTH / 4), c4 = i % (WIDTH / 4);
      long long gk = kb + kk;
      bool valid = gk < k1 && (col0 + c4 * 4) < ncols;
      const float* src = B + (valid ? gk : 0) * ldb + (valid ? col0 + c4 * 4 : 0);
      cp16(Bs + (buf * KC + kk) * WIDTH + c4 * 4, src, valid);
    }
    cp_commit();
  };
  const long long nchunk = (k1 - k0 + KC - 1) / KC;
  if (nchunk <= 0) return;
  load(k0, 0);
  for (long long c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) { load(k0 + (c + 1) * KC, (int)((c + 1) & 1)); cp_wait<1>(); }
    else cp_wait<0>();
    __syncthreads();
    const float* Ab = As + ((c & 1) * TILE_M) * AS;
    const float* Bb = Bs + ((c & 1) * KC) * WIDTH;
#pragma unroll
    for (int k4 = 0; k4 < KC; k4 += 4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i)
        a[i] = *reinterpret_cast<const float4*>(Ab + (ty * TM + i) * AS + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 w0 = *reinterpret_cast<const float4*>(Bb + (k4 + kk) * WIDTH + tx * 4);
        float4 w1 = *reinterpret_cast<const float4*>(Bb + (k4 + kk) * WIDTH + WIDTH / 2 + tx * 4);
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

template <int WIDTH>
constexpr size_t tile_smem() {
  return sizeof(float) * (size_t)(2 * TILE_M * (KC + 4) + 2 * KC * WIDTH);
}

// ---- forward: Y = act([X1 | X2] W + b), grid = (row tiles, column passes of WIDTH)
template <int WIDTH>
__global__ void __launch_bounds__(NTHREADS) dense_fwd_kernel(
    const float* __restrict__ x1, int k1, const float* __restrict__ x2, int k2,
    const float* __restrict__ w, const float* __restrict__ b, long long m, int n, int act,
    float* __restrict__ y) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;
  float* Bs = sm + 2 * TILE_M * (KC + 4);
  constexpr int TM = Tile<WIDTH>::TM;
  const int tid = threadIdx.x, tx = tid % Tile<WIDTH>::TXN, ty = tid / Tile<WIDTH>::TXN;
  const long long row0 = (long long)blockIdx.x * TILE_M;
  const int col0 = blockIdx.y * WIDTH;
  float acc[TM][8];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  gemm_gg<WIDTH>(x1, k1, 0, m, k1, row0, w, n, 0, k1, col0, n, As, Bs, acc, tid);
  if (k2 > 0)
    gemm_gg<WIDTH>(x2, k2, 0, m, k2, row0, w + (size_t)k1 * n, n, 0, k2, col0, n, As, Bs, acc, tid);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long r = row0 + ty * TM + i;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col0 + (j < 4 ? tx * 4 + j : WIDTH / 2 + tx * 4 + (j - 4));
      if (c < n) y[r * n + c] = apply_act(act, acc[i][j] + b[c]);
    }
  }
}

// dz = dy * act'(y)
__global__ void act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                               long long total, int act, float* __restrict__ dz) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) dz[i] = dy[i] * act_grad(act, y[i]);
}

// ---- data grad: dX[m, k] = dZ[m, n] * W^T   (B = W^T given as wt[n][ktot], columns = k)
template <int WIDTH>
__global__ void __launch_bounds__(NTHREADS) dense_dgrad_kernel(
    const float* __restrict__ dz, long long m, int n, const float* __restrict__ wt, int ktot,
    int k1, float* __restrict__ dx1, int k2, float* __restrict__ dx2) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;
  float* Bs = sm + 2 * TILE_M * (KC + 4);
  constexpr int TM = Tile<WIDTH>::TM;
  const int tid = threadIdx.x, tx = tid % Tile<WIDTH>::TXN, ty = tid / Tile<WIDTH>::TXN;
  const long long row0 = (long long)blockIdx.x * TILE_M;
  const int col0 = blockIdx.y * WIDTH;
  float acc[TM][8];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  gemm_gg<WIDTH>(dz, n, 0, m, n, row0, wt, ktot, 0, n, col0, ktot, As, Bs, acc, tid);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long r = row0 + ty * TM + i;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col0 + (j < 4 ? tx * 4 + j : WIDTH / 2 + tx * 4 + (j - 4));
      if (c < k1) { if (dx1) dx1[r * k1 + c] = acc[i][j]; }
      else if (c < k1 + k2) { if (dx2) dx2[r * k2 + (c - k1)] = acc[i][j]; }
    }
  }
}

// ---- weight grad: dW[k, n] += sum_m X[m, k] dZ[m, n]; grid = (k tiles of 64, n passes, m splits)
template <int WIDTH>
__global__ void __launch_bounds__(NTHREADS) dense_wgrad_kernel(
    const float* __restrict__ x, int kx, const float* __restrict__ dz, long long m, int n,
    long long rows_per_split, float* __restrict__ dw /* [kx rows at offset][n] */) {
  extern __shared__ __align__(16) float sm[];
  float* As = sm;
  float* Bs = sm + 2 * TILE_M * (KC + 4);
  constexpr int TM = Tile<WIDTH>::TM;
  const int tid = threadIdx.x, tx = tid % Tile<WIDTH>::TXN, ty = tid / Tile<WIDTH>::TXN;
  const long long krow0 = (long long)blockIdx.x * TILE_M;     // output rows = k index
  const int col0 = blockIdx.y * WIDTH;
  const long long m0 = (long long)blockIdx.z * rows_per_split;
  const long long m1 = m0 + rows_per_split < m ? m0 + rows_per_split : m;
  float acc[TM][8];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // A(i, kk) = X[kk (sample)][krow0 + i]  -> a_trans = 1; reduction index = sample
  gemm_gg<WIDTH>(x, kx, 1, m, kx, krow0, dz, n, m0, m1, col0, n, As, Bs, acc, tid);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long r = krow0 + ty * TM + i;
    if (r >= kx) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = col0 + (j < 4 ? tx * 4 + j : WIDTH / 2 + tx * 4 + (j - 4));
      if (c < n) atomicAdd(dw + r * n + c, acc[i][j]);
    }
  }
}

// db[n] += colsum(dz[m, n])
__global__ void colsum_kernel(const float* __restrict__ dz, long long m, int n,
                              long long rows_per_block, float* __restrict__ db) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < m ? r0 + rows_per_block : m;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += dz[r * n + c];
  atomicAdd(db + c, s);
}

__global__ void transpose_kernel(const float* __restrict__ w, int k, int n, float* __restrict__ wt) {
  __shared__ float t[32][33];
  int kx = blockIdx.x * 32 + threadIdx.y, nx = blockIdx.y * 32 + threadIdx.x;
  if (kx < k && nx < n) t[threadIdx.y][threadIdx.x] = w[(size_t)kx * n + nx];
  __syncthreads();
  int nn = blockIdx.y * 32 + threadIdx.y, kk = blockIdx.x * 32 + threadIdx.x;
  if (nn < n && kk < k) wt[(size_t)nn * k + kk] = t[threadIdx.x][threadIdx.y];
}

// tf.keras.optimizers.Adam(amsgrad=True): m, v, vhat slots; epsilon outside the sqrt
// (trainvali.py:110-127; Keras OptimizerV2 _resource_apply_dense).
__global__ void adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g,
                                    float* __restrict__ m, float* __restrict__ v,
                                    float* __restrict__ vhat, long long count, float lr_t,
                                    float beta1, float beta2, float eps) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float gi = g[i];
  float mi = m[i] + (gi - m[i]) * (1.f - beta1);
  float vi = v[i] + (gi * gi - v[i]) * (1.f - beta2);
  float vh = fmaxf(vhat[i], vi);
  m[i] = mi; v[i] = vi; vhat[i] = vh;
  p[i] = p[i] - lr_t * mi / (sqrtf(vh) + eps);
}

}  // namespace

// nf_train_tc.cu: tcgen05 versions (16-bit operands)
bool nf_dense_tc_supported(int k1, int k2, int n);
size_t nf_dense_tc_fwd_workspace(int k1, int k2, int n);
size_t nf_dense_tc_bwd_workspace(long long m, int k1, int k2, int n, int sm_count);
int nf_dense_tc_fwd(nf_ctx* ctx, const float* x1, int k1, const float* x2, int k2, const float* w,
                    const float* b, long long m, int n, int act, float* y, void* work,
                    int precision, cudaStream_t st);
int nf_dense_tc_bwd(nf_ctx* ctx, const float* x1, int k1, const float* x2, int k2, const float* w,
                    const float* y, const float* dy, long long m, int n, int act, float* dx1,
                    float* dx2, float* dw, float* db, void* work, int precision, cudaStream_t st);

static bool use_tc(int precision, int k1, int k2, int n) {
  return (precision == NF_PREC_F16 || precision == NF_PREC_BF16) && nf_dense_tc_supported(k1, k2, n);
}

extern "C" {

size_t nf_dense_fwd_workspace_bytes(int n, int k1, int k2, int precision) {
  return use_tc(precision, k1, k2, n) ? nf_dense_tc_fwd_workspace(k1, k2, n) : 0;
}

int nf_dense_fwd(nf_ctx* ctx, const float* x1_d, int k1, const float* x2_d, int k2,
                 const float* w_d, const float* b_d, long long m, int n, int act, float* y_d,
                 void* work_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, m >= 0 && n >= 1 && k1 >= 1 && k2 >= 0, "bad sizes");
  if (m == 0) return NF_OK;
  NF_CHECK_ARG(ctx, x1_d && w_d && b_d && y_d && (k2 == 0 || x2_d), "null buffer");
  NF_CHECK_ARG(ctx, n % 4 == 0 || n < 4, "n must be a multiple of 4 (or < 4)");
  NF_CHECK_ARG(ctx, precision == NF_PREC_FP32 || precision == NF_PREC_F16 || precision == NF_PREC_BF16,
               "precision must be NF_PREC_FP32, NF_PREC_F16 or NF_PREC_BF16");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(precision, k1, k2, n)) {
    NF_CHECK_ARG(ctx, work_d, "null workspace (nf_dense_fwd_workspace_bytes)");
    return nf_dense_tc_fwd(ctx, x1_d, k1, x2_d, k2, w_d, b_d, m, n, act, y_d, work_d, precision, st);
  }
  if (n < 4) {
    // tiny heads: pad the weight rows on the fly is not worth a kernel; use W = 128-wide path
    // through a padded copy is the caller's job -> not supported here
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED, "nf_dense_fwd: pad head weights to n = 4");
  }
  const long long tiles = (m + TILE_M - 1) / TILE_M;
  NF_CHECK_ARG(ctx, tiles < 2147483647LL, "too many rows");
  if (n > 128) {
    dim3 grid((unsigned)tiles, (n + 255) / 256);
    size_t smb = tile_smem<256>();
    dense_fwd_kernel<256><<<grid, NTHREADS, smb, st>>>(x1_d, k1, x2_d, k2, w_d, b_d, m, n, act, y_d);
  } else {
    dim3 grid((unsigned)tiles, 1);
    size_t smb = tile_smem<128>();
    dense_fwd_kernel<128><<<grid, NTHREADS, smb, st>>>(x1_d, k1, x2_d, k2, w_d, b_d, m, n, act, y_d);
  }
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

size_t nf_dense_bwd_workspace_bytes(long long m, int n, int k1, int k2, int precision) {
  if (use_tc(precision, k1, k2, n)) return nf_dense_tc_bwd_workspace(m, k1, k2, n, 160);
  size_t dz = (size_t)m * n * sizeof(float);
  size_t wt = (size_t)n * (k1 + k2) * sizeof(float);
  return (dz + 255) / 256 * 256 + (wt + 255) / 256 * 256;
}

int nf_dense_bwd(nf_ctx* ctx, const float* x1_d, int k1, const float* x2_d, int k2,
                 const float* w_d, const float* y_d, const float* dy_d, long long m, int n,
                 int act, float* dx1_d, float* dx2_d, float* dw_d, float* db_d, void* work_d,
                 int precision, void* stream) {
  NF_CHECK_ARG(ctx, m >= 0 && n >= 4 && n % 4 == 0 && k1 >= 1 && k2 >= 0, "bad sizes");
  if (m == 0) return NF_OK;
  NF_CHECK_ARG(ctx, x1_d && w_d && y_d && dy_d && work_d && (k2 == 0 || x2_d), "null buffer");
  NF_CHECK_ARG(ctx, k1 % 4 == 0 && k2 % 4 == 0, "k1, k2 must be multiples of 4");
  NF_CHECK_ARG(ctx, precision == NF_PREC_FP32 || precision == NF_PREC_F16 || precision == NF_PREC_BF16,
               "precision must be NF_PREC_FP32, NF_PREC_F16 or NF_PREC_BF16");
  cudaStream_t st = (cudaStream_t)stream;
  if (use_tc(precision, k1, k2, n)) {
    NF_CHECK_ARG(ctx, ctx->sm_count <= 160, "unexpected SM count");
    return nf_dense_tc_bwd(ctx, x1_d, k1, x2_d, k2, w_d, y_d, dy_d, m, n, act, dx1_d, dx2_d, dw_d,
                           db_d, work_d, precision, st);
  }
  const int ktot = k1 + k2;
  float* dz = reinterpret_cast<float*>(work_d);
  float* wt = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(work_d) +
                                       ((size_t)m * n * sizeof(float) + 255) / 256 * 256);
  const long long total = m * n;
  act_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(y_d, dy_d, total, act, dz);
  NF_LAUNCH_CHECK(ctx);
  const long long tiles = (m + TILE_M - 1) / TILE_M;
  if (dx1_d || dx2_d) {
    dim3 tg((ktot + 31) / 32, (n + 31) / 32);
    transpose_kernel<<<tg, dim3(32, 32), 0, st>>>(w_d, ktot, n, wt);
    NF_LAUNCH_CHECK(ctx);
    if (ktot > 128) {
      dim3 grid((unsigned)tiles, (ktot + 255) / 256);
      dense_dgrad_kernel<256><<<grid, NTHREADS, tile_smem<256>(), st>>>(dz, m, n, wt, ktot, k1, dx1_d, k2, dx2_d);
    } else {
      dim3 grid((unsigned)tiles, 1);
      dense_dgrad_kernel<128><<<grid, NTHREADS, tile_smem<128>(), st>>>(dz, m, n, wt, ktot, k1, dx1_d, k2, dx2_d);
    }
    NF_LAUNCH_CHECK(ctx);
  }
  if (dw_d) {
    const long long rows_per_split = 4096;
    const unsigned splits = (unsigned)((m + rows_per_split - 1) / rows_per_split);
    for (int seg = 0; seg < 2; ++seg) {
      const float* x = seg == 0 ? x1_d : x2_d;
      const int kx = seg == 0 ? k1 : k2;
      if (kx == 0) continue;
      float* dws = dw_d + (size_t)(seg == 0 ? 0 : k1) * n;
      if (n > 128) {
        dim3 grid((kx + TILE_M - 1) / TILE_M, (n + 255) / 256, splits);
        dense_wgrad_kernel<256><<<grid, NTHREADS, tile_smem<256>(), st>>>(x, kx, dz, m, n, rows_per_split, dws);
      } else {
        dim3 grid((kx + TILE_M - 1) / TILE_M, 1, splits);
        dense_wgrad_kernel<128><<<grid, NTHREADS, tile_smem<128>(), st>>>(x, kx, dz, m, n, rows_per_split, dws);
      }
      NF_LAUNCH_CHECK(ctx);
    }
  }
  if (db_d) {
    const long long rpb = 2048;
    dim3 grid((unsigned)((m + rpb - 1) / rpb), (n + 127) / 128);
    colsum_kernel<<<grid, 128, 0, st>>>(dz, m, n, rpb, db_d);
    NF_LAUNCH_CHECK(ctx);
  }
  return NF_OK;
}

int nf_adam_amsgrad_step(nf_ctx* ctx, float* param_d, const float* grad_d, float* m_d, float* v_d,
                         float* vhat_d, long long count, float lr, float beta1, float beta2,
                         float eps, long long step, void* stream) {
  NF_CHECK_ARG(ctx, count >= 0 && step >= 1, "bad count / step");
  if (count == 0) return NF_OK;
  NF_CHECK_ARG(ctx, param_d && grad_d && m_d && v_d && vhat_d, "null buffer");
  // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)   (Keras Adam._prepare_local)
  const double b1t = pow((double)beta1, (double)step), b2t = pow((double)beta2, (double)step);
  const float lr_t = (float)((double)lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  adam_amsgrad_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      param_d, grad_d, m_d, v_d, vhat_d, count, lr_t, beta1, beta2, eps);
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}

}  // extern "C"
