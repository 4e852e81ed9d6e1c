// Fused skip-MLP forward on the FP32 CUDA cores (NF_PREC_FP32).
//
// One CTA = 64 rows; the whole Dense chain (nerfactor/networks/mlp.py:39-50) runs
// on-chip: the embedded input stays in shared memory for the skip concat,
// activations ping-pong between two shared buffers, weights stream from L2 in
// 16-row chunks (cp.async double buffer).  The input prologue builds each row
// straight from the caller's geometry (no [N*L, 90] tensor as in shape.py:219-233):
//   POINT  embed(xyz)                          shape.py:196-211, nerfactor.py:377-411
//   LVIS   embed(xyz) ++ embed(l2n(lxyz-xyz))  shape.py:128-135, 213-237
//   BRDF   z ++ embed(rusink)                  nerfactor.py:413-452
//   SIGMA  embed(o + z d), bbox mask           geometry_from_nerf.py:322-350
// This is the exact-fp32 path: parity reference on the device, and the product
// path for the three tiny per-point networks.
#include "nf_common.cuh"

namespace {

constexpr int TILE_M = 64;
constexpr int KC = 16;           // weight rows per chunk
constexpr int X_STRIDE = 100;    // >= in_pad (96), multiple of 4
constexpr int NTHREADS = 256;

struct SimtParams {
  const uint8_t* blob;           // packed weights (device)
  size_t off_w[9], off_b[9];
  int kind, in_dim, in_pad, depth, skip_at, out_dim, out_act;
  int n_freqs_a, n_freqs_b, z_dim;
  long long n_rows;              // total MLP rows
  int per;                       // rows per ray (L or S; 1 for POINT)
  float xyz_scale;
  const float* xyz;              // POINT/LVIS/BRDF: [n,3]; SIGMA: rayo [n,3]
  const float* aux0;             // LVIS/BRDF: lxyz [L,3]; SIGMA: rayd [n,3]
  const float* aux1;             // BRDF: normal [n,3]; SIGMA: z [n,S]
  const float* aux2;             // BRDF: cam [n,3]
  const float* aux3;             // BRDF: z latent [n,z_dim]
  float bbox[6];
  int use_bbox;
  float* out;                    // [n_rows, out_dim]
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// writes embed(v) for component c of a 3-vector into row[base ...] (embedder.py:46-47)
__device__ __forceinline__ void embed_comp(float* row, int base, int c, float v, int n_freqs) {
  row[base + c] = v;
  float f = 1.f;
  for (int k = 0; k < n_freqs; ++k) {
    float s, co;
    sincosf(v * f, &s, &co);
    row[base + 3 + 6 * k + c] = s;
    row[base + 3 + 6 * k + 3 + c] = co;
    f *= 2.f;
  }
}

template <int WIDTH>
__global__ void __launch_bounds__(NTHREADS) mlp_simt_kernel(const SimtParams p) {
  constexpr int H_STRIDE = WIDTH + 4;
  constexpr int TXN = WIDTH / 8;          // threads along N (16 / 32)
  constexpr int TYN = NTHREADS / TXN;     // threads along M (16 / 8)
  constexpr int TM = TILE_M / TYN;        // rows per thread (4 / 8)
  extern __shared__ __align__(16) float smem[];
  float* X0 = smem;                              // [TILE_M][X_STRIDE]
  float* H0 = X0 + TILE_M * X_STRIDE;            // [TILE_M][H_STRIDE]
  float* H1 = H0 + TILE_M * H_STRIDE;
  float* Wc = H1 + TILE_M * H_STRIDE;            // [2][KC][WIDTH]
  float* rus = Wc + 2 * KC * WIDTH;              // [TILE_M][4] rusink scratch + mask
  __shared__ float s_mask[TILE_M];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const long long row0 = (long long)blockIdx.x * TILE_M;

  // ------------------------------------------------------------ prologue
  {
    const int r = tid >> 2, sub = tid & 3;
    const long long g = row0 + r;
    float* xr = X0 + r * X_STRIDE;
    const bool live = g < p.n_rows;
    if (sub == 3) {
      for (int k = p.in_dim; k < p.in_pad; ++k) xr[k] = 0.f;
      if (!live) for (int k = 0; k < p.in_dim; ++k) xr[k] = 0.f;
      if (p.kind != NF_MLP_BRDF && p.kind != NF_MLP_SIGMA) s_mask[r] = 1.f;
    }
    long long ray = live ? g / p.per : 0;
    int sub_i = live ? (int)(g % p.per) : 0;
    if (p.kind == NF_MLP_POINT) {
      if (live && sub < 3) embed_comp(xr, 0, sub, p.xyz[ray * 3 + sub] * p.xyz_scale, p.n_freqs_a);
    } else if (p.kind == NF_MLP_LVIS) {
      if (live && sub < 3) {
        f3 pt = ld3(p.xyz + ray * 3);
        f3 d = l2n(ld3(p.aux0 + sub_i * 3) - pt, 1e-6f);          // shape.py:129-131
        float pc = sub == 0 ? pt.x : (sub == 1 ? pt.y : pt.z);
        float dc = sub == 0 ? d.x : (sub == 1 ? d.y : d.z);
        embed_comp(xr, 0, sub, pc * p.xyz_scale, p.n_freqs_a);
        embed_comp(xr, 3 * (1 + 2 * p.n_freqs_a), sub, dc, p.n_freqs_b);
      }
    } else if (p.kind == NF_MLP_SIGMA) {
      f3 o = ld3(p.xyz + ray * 3), d = ld3(p.aux0 + ray * 3);
      float z = live ? p.aux1[g] : 0.f;
      // pts = rayo + rayd * z (gfn.py:264): multiply then add, like the TF / oracle ops
      f3 pt = mk3(__fadd_rn(o.x, __fmul_rn(d.x, z)), __fadd_rn(o.y, __fmul_rn(d.y, z)),
                  __fadd_rn(o.z, __fmul_rn(d.z, z)));
      if (live && sub < 3) {
        float pc = sub == 0 ? pt.x : (sub == 1 ? pt.y : pt.z);
        embed_comp(xr, 0, sub, pc, p.n_freqs_a);
      }
      if (sub == 3) {
        bool in = true;
        if (p.use_bbox)
          in = pt.x >= p.bbox[0] && pt.x <= p.bbox[1] && pt.y >= p.bbox[2] &&
               pt.y <= p.bbox[3] && pt.z >= p.bbox[4] && pt.z <= p.bbox[5];
        s_mask[r] = in ? 1.f : 0.f;
      }
    } else {  // BRDF
      if (sub == 0) {
        f3 pt = ld3(p.xyz + ray * 3);
        f3 t, b, n;
        world2local_dev(ld3(p.aux1 + ray * 3), t, b, n);                 // geom.py:119-149
        f3 l = l2n(ld3(p.aux0 + sub_i * 3) - pt, 1e-6f);                  // shape.py:128-135
        f3 v = l2n(ld3(p.aux2 + ray * 3) - pt, 1e-6f);                    // shape.py:137-144
        f3 ll = mk3(dot3(t, l), dot3(b, l), dot3(n, l));                  // nerfactor.py:418-419
        f3 vl = mk3(dot3(t, v), dot3(b, v), dot3(n, v));
        f3 rs = dir2rusink_dev(ll, vl);
        rus[r * 4 + 0] = rs.x; rus[r * 4 + 1] = rs.y; rus[r * 4 + 2] = rs.z;
        s_mask[r] = (live && ll.z > 0.f) ? 1.f : 0.f;                     // nerfactor.py:429-432
      }
      __syncwarp();
      if (live) {
        if (sub < 3) embed_comp(xr, p.z_dim, sub, rus[r * 4 + sub], p.n_freqs_a);
        if (sub == 3) for (int k = 0; k < p.z_dim; ++k) xr[k] = p.aux3[ray * p.z_dim + k];
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------ Dense chain
  float* Hin = nullptr;
  float* Hout = H0;
  for (int l = 0; l < p.depth; ++l) {
    const float* Wg = reinterpret_cast<const float*>(p.blob + p.off_w[l]);
    const float* bg = reinterpret_cast<const float*>(p.blob + p.off_b[l]);
    // K segments: (source buffer, stride, real K, weight-row offset)
    int nseg = 1;
    const float* segA[2]; int segS[2], segK[2], segW[2];
    if (l == 0) { segA[0] = X0; segS[0] = X_STRIDE; segK[0] = p.in_dim; segW[0] = 0; }
    else {
      segA[0] = Hin; segS[0] = H_STRIDE; segK[0] = WIDTH; segW[0] = 0;
      if (l == p.skip_at + 1) { nseg = 2; segA[1] = X0; segS[1] = X_STRIDE; segK[1] = p.in_dim; segW[1] = WIDTH; }
    }
    float acc[TM][8];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int s = 0; s < nseg; ++s) {
      const int kreal = segK[s];
      const int nchunk = (kreal + KC - 1) / KC;
      const float* A = segA[s];
      const int as = segS[s];
      auto load_chunk = [&](int c, int buf) {
        // KC x WIDTH floats, 16 B per cp.async
        for (int i = tid; i < KC * WIDTH / 4; i += NTHREADS) {
          int kk = i / (WIDTH / 4), c4 = i % (WIDTH / 4);
          int krow = c * KC + kk;
          bool valid = krow < kreal;
          const float* src = Wg + (size_t)(segW[s] + (valid ? krow : 0)) * WIDTH + c4 * 4;
          cp_async16(Wc + buf * KC * WIDTH + kk * WIDTH + c4 * 4, src, valid);
        }
        cp_async_commit();
      };
      load_chunk(0, 0);
      for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) { load_chunk(c + 1, (c + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
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
    // bias + ReLU -> Hout
    float4 b0 = *reinterpret_cast<const float4*>(bg + tx * 4);
    float4 b1 = *reinterpret_cast<const float4*>(bg + WIDTH / 2 + tx * 4);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      float4 o0, o1;
      o0.x = fmaxf(acc[i][0] + b0.x, 0.f); o0.y = fmaxf(acc[i][1] + b0.y, 0.f);
      o0.z = fmaxf(acc[i][2] + b0.z, 0.f); o0.w = fmaxf(acc[i][3] + b0.w, 0.f);
      o1.x = fmaxf(acc[i][4] + b1.x, 0.f); o1.y = fmaxf(acc[i][5] + b1.y, 0.f);
      o1.z = fmaxf(acc[i][6] + b1.z, 0.f); o1.w = fmaxf(acc[i][7] + b1.w, 0.f);
      *reinterpret_cast<float4*>(Hout + (ty * TM + i) * H_STRIDE + tx * 4) = o0;
      *reinterpret_cast<float4*>(Hout + (ty * TM + i) * H_STRIDE + WIDTH / 2 + tx * 4) = o1;
    }
    __syncthreads();
    Hin = Hout;
    Hout = (Hout == H0) ? H1 : H0;
  }

  // ------------------------------------------------------------ head
  {
    const float* Wo = reinterpret_cast<const float*>(p.blob + p.off_w[p.depth]);  // [WIDTH][out_dim]
    const float* bo = reinterpret_cast<const float*>(p.blob + p.off_b[p.depth]);
    const int r = tid >> 2, q = tid & 3;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = q; c < WIDTH; c += 4) {
      float h = Hin[r * H_STRIDE + c];
      for (int j = 0; j < p.out_dim; ++j) o[j] = fmaf(h, Wo[c * p.out_dim + j], o[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] += __shfl_xor_sync(0xffffffffu, o[j], 1);
      o[j] += __shfl_xor_sync(0xffffffffu, o[j], 2);
    }
    const long long g = row0 + r;
    if (g < p.n_rows && q < p.out_dim) {
      float v = o[q] + bo[q];
      v = apply_act(p.out_act, v);
      if (p.kind == NF_MLP_SIGMA) v = fmaxf(v, 0.f);   // tf.nn.relu, gfn.py:340
      p.out[g * p.out_dim + q] = v * s_mask[r];
    }
  }
}

template <int WIDTH>
size_t simt_smem_bytes() {
  return sizeof(float) * (size_t)(TILE_M * X_STRIDE + 2 * TILE_M * (WIDTH + 4) +
                                  2 * KC * WIDTH + TILE_M * 4);
}

}  // namespace

int nf_simt_launch(nf_ctx* ctx, const nf_mlp* m, long long n_rows, int per, float xyz_scale,
                   const float* xyz, const float* a0, const float* a1, const float* a2,
                   const float* a3, const float* bbox_host, float* out, cudaStream_t st) {
  if (n_rows == 0) return NF_OK;
  NF_CHECK_ARG(ctx, m && m->dev, "network not uploaded (call nf_mlp_upload first)");
  SimtParams p;
  memset(&p, 0, sizeof(p));
  p.blob = (const uint8_t*)m->dev;
  for (int l = 0; l <= m->d.depth; ++l) { p.off_w[l] = m->off_w32[l]; p.off_b[l] = m->off_b32[l]; }
  p.kind = m->d.kind; p.in_dim = m->d.in_dim; p.in_pad = m->in_pad; p.depth = m->d.depth;
  p.skip_at = m->d.skip_at; p.out_dim = m->d.out_dim; p.out_act = m->d.out_act;
  p.n_freqs_a = m->d.n_freqs_a; p.n_freqs_b = m->d.n_freqs_b; p.z_dim = m->d.z_dim;
  p.n_rows = n_rows; p.per = per; p.xyz_scale = xyz_scale;
  p.xyz = xyz; p.aux0 = a0; p.aux1 = a1; p.aux2 = a2; p.aux3 = a3; p.out = out;
  if (bbox_host) { memcpy(p.bbox, bbox_host, sizeof(p.bbox)); p.use_bbox = 1; }
  long long blocks = (n_rows + TILE_M - 1) / TILE_M;
  NF_CHECK_ARG(ctx, blocks < 2147483647LL, "too many rows for one launch");
  if (m->d.width == 128) {
    size_t sm = simt_smem_bytes<128>();
    NF_CUDA(ctx, cudaFuncSetAttribute(mlp_simt_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    mlp_simt_kernel<128><<<(unsigned)blocks, NTHREADS, sm, st>>>(p);
  } else {
    size_t sm = simt_smem_bytes<256>();
    NF_CUDA(ctx, cudaFuncSetAttribute(mlp_simt_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    mlp_simt_kernel<256><<<(unsigned)blocks, NTHREADS, sm, st>>>(p);
  }
  NF_LAUNCH_CHECK(ctx);
  return NF_OK;
}
