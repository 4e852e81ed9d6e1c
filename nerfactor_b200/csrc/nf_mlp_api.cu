// C-ABI entry points of the network forward passes: argument checking and
// dispatch between the FP32 CUDA-core kernels and the tcgen05 kernels.
#include "nf_common.cuh"

int nf_simt_launch(nf_ctx* ctx, const nf_mlp* m, long long n_rows, int per, float xyz_scale,
                   const float* xyz, const float* a0, const float* a1, const float* a2,
                   const float* a3, const float* bbox_host, float* out, cudaStream_t st);
// nf_point_tc.cu
int nf_tc_point_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, int n, float xyz_scale,
                       float* out, cudaStream_t st);
// nf_mlp_tc.cu
int nf_tc_lvis_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, int n, float xyz_scale,
                      const float* lxyz, int L, float* lvis, int precision, cudaStream_t st,
                      const float* xyz_dir = nullptr, const float* cull_normal = nullptr);
int nf_tc_sigma_launch(nf_ctx* ctx, const nf_mlp* m, const float* rayo, const float* rayd,
                       const float* z, int n_rays, int S, const float* bbox_host, float* sigma,
                       int precision, cudaStream_t st);
int nf_tc_brdf_launch(nf_ctx* ctx, const nf_mlp* m, const float* xyz, const float* normal,
                      const float* cam, const float* z, int n, const float* lxyz, int L,
                      float* spec, int precision, cudaStream_t st);

extern "C" {

int nf_point_mlp_fwd(nf_ctx* ctx, const nf_mlp* mlp, const float* xyz_d, int n, float xyz_scale,
                     float* out_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, mlp && n >= 0, "bad argument");
  NF_CHECK_ARG(ctx, mlp->d.kind == NF_MLP_POINT, "network is not NF_MLP_POINT");
  if (n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, xyz_d && out_d, "null buffer");
  if (precision == NF_PREC_F16X3)
    return nf_tc_point_launch(ctx, mlp, xyz_d, n, xyz_scale, out_d, (cudaStream_t)stream);
  if (precision != NF_PREC_FP32)
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "nf_point_mlp_fwd: the per-point networks need NF_PREC_FP32 or "
                        "NF_PREC_F16X3 (plain 16-bit operands are not accurate enough)");
  return nf_simt_launch(ctx, mlp, n, 1, xyz_scale, xyz_d, nullptr, nullptr, nullptr, nullptr,
                        nullptr, out_d, (cudaStream_t)stream);
}

int nf_lvis_fwd(nf_ctx* ctx, const nf_mlp* mlp, const float* xyz_d, int n, float xyz_scale,
                const float* lxyz_d, int n_lights, float* lvis_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, mlp && lxyz_d && n >= 0 && n_lights > 0, "bad argument");
  NF_CHECK_ARG(ctx, mlp->d.kind == NF_MLP_LVIS && mlp->d.out_dim == 1, "network is not NF_MLP_LVIS");
  if (n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, xyz_d && lvis_d, "null buffer");
  if (precision == NF_PREC_FP32)
    return nf_simt_launch(ctx, mlp, (long long)n * n_lights, n_lights, xyz_scale, xyz_d, lxyz_d,
                          nullptr, nullptr, nullptr, nullptr, lvis_d, (cudaStream_t)stream);
  return nf_tc_lvis_launch(ctx, mlp, xyz_d, n, xyz_scale, lxyz_d, n_lights, lvis_d, precision,
                           (cudaStream_t)stream);
}

int nf_lvis_dirs_fwd(nf_ctx* ctx, const nf_mlp* mlp, const float* xyz_d, const float* xyz_dir_d,
                     int n, float xyz_scale, const float* lxyz_d, int n_lights, float* lvis_d,
                     int precision, void* stream) {
  NF_CHECK_ARG(ctx, mlp && lxyz_d && n >= 0 && n_lights > 0, "bad argument");
  NF_CHECK_ARG(ctx, mlp->d.kind == NF_MLP_LVIS && mlp->d.out_dim == 1, "network is not NF_MLP_LVIS");
  if (n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, xyz_d && xyz_dir_d && lvis_d, "null buffer");
  if (precision != NF_PREC_F16 && precision != NF_PREC_BF16)
    return nf_set_error(ctx, NF_ERR_UNSUPPORTED,
                        "nf_lvis_dirs_fwd: tcgen05 only (NF_PREC_F16 / NF_PREC_BF16)");
  return nf_tc_lvis_launch(ctx, mlp, xyz_d, n, xyz_scale, lxyz_d, n_lights, lvis_d, precision,
                           (cudaStream_t)stream, xyz_dir_d);
}

int nf_brdf_learned_fwd(nf_ctx* ctx, const nf_mlp* mlp, const float* xyz_d, const float* normal_d,
                        const float* cam_d, const float* z_d, int n, const float* lxyz_d,
                        int n_lights, float* spec_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, mlp && lxyz_d, "null argument");
  NF_CHECK_ARG(ctx, n >= 0 && n_lights > 0, "bad sizes");
  NF_CHECK_ARG(ctx, mlp->d.kind == NF_MLP_BRDF && mlp->d.out_dim == 1, "network is not NF_MLP_BRDF");
  if (n == 0) return NF_OK;
  NF_CHECK_ARG(ctx, xyz_d && normal_d && cam_d && z_d && spec_d, "null buffer");
  if (precision == NF_PREC_FP32)
    return nf_simt_launch(ctx, mlp, (long long)n * n_lights, n_lights, 1.f, xyz_d, lxyz_d,
                          normal_d, cam_d, z_d, nullptr, spec_d, (cudaStream_t)stream);
  return nf_tc_brdf_launch(ctx, mlp, xyz_d, normal_d, cam_d, z_d, n, lxyz_d, n_lights, spec_d,
                           precision, (cudaStream_t)stream);
}

int nf_sigma_fwd(nf_ctx* ctx, const nf_mlp* mlp, const float* rayo_d, const float* rayd_d,
                 const float* z_d, int n_rays, int n_samples, const float* bbox_host,
                 float* sigma_d, int precision, void* stream) {
  NF_CHECK_ARG(ctx, mlp, "null argument");
  NF_CHECK_ARG(ctx, n_rays >= 0 && n_samples > 0, "bad sizes");
  NF_CHECK_ARG(ctx, mlp->d.kind == NF_MLP_SIGMA && mlp->d.out_dim == 1, "network is not NF_MLP_SIGMA");
  if (n_rays == 0) return NF_OK;
  NF_CHECK_ARG(ctx, rayo_d && rayd_d && z_d && sigma_d, "null buffer");
  if (precision == NF_PREC_FP32)
    return nf_simt_launch(ctx, mlp, (long long)n_rays * n_samples, n_samples, 1.f, rayo_d, rayd_d,
                          z_d, nullptr, nullptr, bbox_host, sigma_d, (cudaStream_t)stream);
  return nf_tc_sigma_launch(ctx, mlp, rayo_d, rayd_d, z_d, n_rays, n_samples, bbox_host, sigma_d,
                            precision, (cudaStream_t)stream);
}

}  // extern "C"
