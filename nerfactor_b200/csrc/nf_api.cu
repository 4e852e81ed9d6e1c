// Context, error reporting and host-side weight packing.
#include <stdarg.h>

#include "nf_common.cuh"

int nf_tc_pack(nf_mlp* m);  // nf_mlp_tc.cu: appends the tcgen05 operand images
int nf_sigma_grad_pack(nf_mlp* m);  // nf_sigma_grad.cu: appends W^T for the input gradient
int nf_sigma_grad_tc_pack(nf_mlp* m);  // nf_sigma_grad_tc.cu: backward tcgen05 operand images

int nf_set_error(nf_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->last_error = buf;
  return code;
}

extern "C" {

int nf_version(void) { return 100; }

int nf_ctx_create(nf_ctx** out, int device) {
  if (!out) return NF_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return NF_ERR_NO_DEVICE;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return NF_ERR_NO_DEVICE;
  }
  if (device >= count) return NF_ERR_INVALID_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NF_ERR_CUDA;
  if (prop.major != 10) return NF_ERR_NO_DEVICE;  // sm_100a code only
  nf_ctx* c = new nf_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  *out = c;
  return NF_OK;
}

int nf_ctx_destroy(nf_ctx* ctx) {
  delete ctx;
  return NF_OK;
}

const char* nf_last_error_string(const nf_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : "null context";
}

int nf_ctx_sm_count(const nf_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

// K (input size) of trunk layer l as Keras infers it (mlp.py:39-50)
static int layer_k(const nf_mlp_desc& d, int l) {
  if (l == 0) return d.in_dim;
  if (l == d.skip_at + 1) return d.width + d.in_dim;
  return d.width;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int nf_mlp_create(nf_ctx* ctx, const nf_mlp_desc* desc, nf_mlp** out) {
  NF_CHECK_ARG(ctx, desc && out, "null argument");
  *out = nullptr;
  const nf_mlp_desc& d = *desc;
  NF_CHECK_ARG(ctx, d.kind >= NF_MLP_POINT && d.kind <= NF_MLP_SIGMA, "bad kind");
  NF_CHECK_ARG(ctx, d.width == 128 || d.width == 256, "width must be 128 or 256");
  NF_CHECK_ARG(ctx, d.depth >= 2 && d.depth <= 8, "depth must be in [2, 8]");
  NF_CHECK_ARG(ctx, d.skip_at >= 0 && d.skip_at < d.depth - 1, "skip_at out of range");
  NF_CHECK_ARG(ctx, d.out_dim >= 1 && d.out_dim <= 4, "out_dim must be in [1, 4]");
  NF_CHECK_ARG(ctx, d.in_dim >= 1 && d.in_dim <= 96, "in_dim must be in [1, 96]");
  NF_CHECK_ARG(ctx, d.W && d.b, "null weights");
  int expect_in = 0;
  if (d.kind == NF_MLP_POINT || d.kind == NF_MLP_SIGMA) expect_in = 3 * (1 + 2 * d.n_freqs_a);
  if (d.kind == NF_MLP_LVIS) expect_in = 3 * (1 + 2 * d.n_freqs_a) + 3 * (1 + 2 * d.n_freqs_b);
  if (d.kind == NF_MLP_BRDF) expect_in = d.z_dim + 3 * (1 + 2 * d.n_freqs_a);
  NF_CHECK_ARG(ctx, expect_in == d.in_dim, "in_dim does not match the embedding spec");
  if (d.kind == NF_MLP_BRDF) NF_CHECK_ARG(ctx, d.z_dim >= 1 && d.z_dim <= 8, "bad z_dim");

  nf_mlp* m = new nf_mlp();
  m->d = d;
  m->in_pad = nf_round_up(d.in_dim, 16);
  size_t off = 0;
  m->off_w32.resize(d.depth + 1);
  m->off_b32.resize(d.depth + 1);
  for (int l = 0; l <= d.depth; ++l) {
    int k = l < d.depth ? layer_k(d, l) : d.width;
    int n = l < d.depth ? d.width : d.out_dim;
    m->off_w32[l] = off;
    off = align_up(off + (size_t)k * n * sizeof(float), 256);
    m->off_b32[l] = off;
    off = align_up(off + (size_t)n * sizeof(float), 256);
  }
  m->blob.assign(off, 0);
  for (int l = 0; l <= d.depth; ++l) {
    int k = l < d.depth ? layer_k(d, l) : d.width;
    int n = l < d.depth ? d.width : d.out_dim;
    if (!d.W[l] || !d.b[l]) {
      delete m;
      return nf_set_error(ctx, NF_ERR_INVALID_ARG, "nf_mlp_create: null W/b at layer %d", l);
    }
    memcpy(m->blob.data() + m->off_w32[l], d.W[l], (size_t)k * n * sizeof(float));
    memcpy(m->blob.data() + m->off_b32[l], d.b[l], (size_t)n * sizeof(float));
  }
  int rc = nf_tc_pack(m);
  if (rc == NF_OK) rc = nf_sigma_grad_pack(m);
  if (rc == NF_OK) rc = nf_sigma_grad_tc_pack(m);
  if (rc != NF_OK) {
    delete m;
    return nf_set_error(ctx, rc, "nf_mlp_create: tensor-core packing failed");
  }
  m->d.W = nullptr;
  m->d.b = nullptr;
  *out = m;
  return NF_OK;
}

int nf_mlp_destroy(nf_mlp* mlp) {
  delete mlp;
  return NF_OK;
}

size_t nf_mlp_device_bytes(const nf_mlp* mlp) { return mlp ? mlp->blob.size() : 0; }

int nf_mlp_upload(nf_ctx* ctx, nf_mlp* mlp, void* dst_d, void* stream) {
  NF_CHECK_ARG(ctx, mlp && dst_d, "null argument");
  NF_CHECK_ARG(ctx, ((uintptr_t)dst_d & 255) == 0, "device buffer must be 256-byte aligned");
  NF_CUDA(ctx, cudaMemcpyAsync(dst_d, mlp->blob.data(), mlp->blob.size(),
                               cudaMemcpyHostToDevice, (cudaStream_t)stream));
  mlp->dev = dst_d;
  return NF_OK;
}

}  // extern "C"
