// deform_conv2d.cu — deformable convolution forward (DCNv1/v2), sm_100a.
//
// Reference: csrc/ops/cuda/deform_conv2d_kernel.cu:97-209 (bilinear + deformable_im2col),
// :1035-1255 (host: materialised `columns` buffer + per-group cuBLAS addmm + transpose/copy/bias);
// CPU twin csrc/ops/cpu/deform_conv2d_kernel.cpp:95-209,921-1151.
//
// Design (not a port): the im2col matrix is never written to HBM.  The op is an
// implicit GEMM  out[oc, pix] = sum_k W[oc, k] * col[k, pix],  k = (ci, tap):
//   * SIMT kernel (this file; fp32 / fp16 / bf16 storage, fp32 accumulate): a CTA owns
//     a 128(oc) x 64(pix) output tile of one image.  Per offset group it builds a
//     sampling table in shared memory ONCE — for each (tap, pixel): 4 clamped corner
//     offsets + 4 bilinear weights (zeroed out of bounds, pre-multiplied by the
//     modulation mask) — and reuses it for every input channel of that group, so a
//     col element costs 4 loads + 4 FMAs.  K is walked in slabs of 16; A (weights) and
//     B (sampled columns) slabs live in shared memory, each thread accumulates an 8x4
//     register tile.  Bias is fused into the epilogue; output is written once, NCHW.
//   * tcgen05 kernel (deform_conv2d_tc.cu): same decomposition with the B slab written
//     as a swizzled bf16 K-major tile and the contraction on the 5th-gen tensor cores.
#include "common.cuh"
#include "dcn_params.h"

namespace vb200 {


namespace {

constexpr int BM = 128, BN = 64, BK = 16, DCN_THREADS = 256;
constexpr int BMP = BM + 4;   // padded A-slab row (bank spread for the transposing store)

struct SampleEnt { int o[4]; float w[4]; };   // 32 B

template <typename T>
__global__ void __launch_bounds__(DCN_THREADS)
deform_conv2d_simt_kernel(const T* __restrict__ input, const T* __restrict__ weight, const T* __restrict__ offset,
                          const T* __restrict__ mask, const T* __restrict__ bias, T* __restrict__ out, DcnParams p) {
  extern __shared__ __align__(16) unsigned char dsm[];
  const int KK = p.kh * p.kw;
  SampleEnt* tab = reinterpret_cast<SampleEnt*>(dsm);                 // [KK][BN]
  float* As = reinterpret_cast<float*>(tab + (size_t)KK * BN);        // [BK][BM]
  float* Bs = As + BK * BMP;                                          // [BK][BN]

  const int tid = threadIdx.x;
  const int HWo = p.out_h * p.out_w;
  const int pix0 = blockIdx.x * BN;
  const int cout_g = p.c_out / p.groups, cin_g = p.c_in / p.groups;
  const int m_tiles = ceil_div(cout_g, BM);
  const int g = blockIdx.y / m_tiles;
  const int oc0 = (blockIdx.y % m_tiles) * BM;          // within group
  const int b = blockIdx.z;
  const int c_per_off = p.c_in / p.offset_groups;
  const int Kg = cin_g * KK;                            // weight row length for this group

  const int tm = tid / 16, tn = tid % 16;               // 16 x 16 threads; micro-tile 8 (m) x 4 (n)
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const T* __restrict__ in_b = input + (int64_t)b * p.c_in * p.in_h * p.in_w;
  const int ci_lo = g * cin_g, ci_hi = ci_lo + cin_g;   // channels of this weight group
  const int og_lo = ci_lo / c_per_off, og_hi = (ci_hi - 1) / c_per_off;

  for (int og = og_lo; og <= og_hi; ++og) {
    // ---- sampling table for (offset group og, this pixel tile) ----
    __syncthreads();
    const T* __restrict__ off_b = offset + ((int64_t)b * p.offset_groups + og) * 2 * KK * HWo;
    const T* __restrict__ msk_b = p.use_mask ? mask + ((int64_t)b * p.offset_groups + og) * KK * HWo : nullptr;
    for (int e = tid; e < KK * BN; e += DCN_THREADS) {
      const int tap = e / BN, px = e - tap * BN;
      const int pix = pix0 + px;
      SampleEnt se;
#pragma unroll
      for (int q = 0; q < 4; ++q) { se.o[q] = 0; se.w[q] = 0.f; }
      if (pix < HWo) {
        const int oy = pix / p.out_w, ox = pix - oy * p.out_w;
        const int i = tap / p.kw, j = tap - i * p.kw;
        const float oh = to_acc(off_b[(int64_t)(2 * tap) * HWo + pix]);
        const float ow = to_acc(off_b[(int64_t)(2 * tap + 1) * HWo + pix]);
        const float mv = p.use_mask ? to_acc(msk_b[(int64_t)tap * HWo + pix]) : 1.f;
        const float y = add_rn((float)(oy * p.stride_h - p.pad_h + i * p.dil_h), oh);
        const float x = add_rn((float)(ox * p.stride_w - p.pad_w + j * p.dil_w), ow);
        if (!(y <= -1.f || (float)p.in_h <= y || x <= -1.f || (float)p.in_w <= x)) {
          const int hl = (int)floorf(y), wl = (int)floorf(x);
          const int hh_i = hl + 1, wh_i = wl + 1;
          const float lh = sub_rn(y, (float)hl), lw = sub_rn(x, (float)wl);
          const float hh = sub_rn(1.f, lh), hw = sub_rn(1.f, lw);
          const bool t0 = hl >= 0, t1 = hh_i <= p.in_h - 1, l0 = wl >= 0, l1 = wh_i <= p.in_w - 1;
          const int hlc = max(hl, 0), hhc = min(hh_i, p.in_h - 1), wlc = max(wl, 0), whc = min(wh_i, p.in_w - 1);
          se.o[0] = hlc * p.in_w + wlc; se.w[0] = (t0 && l0) ? mv * (hh * hw) : 0.f;
          se.o[1] = hlc * p.in_w + whc; se.w[1] = (t0 && l1) ? mv * (hh * lw) : 0.f;
          se.o[2] = hhc * p.in_w + wlc; se.w[2] = (t1 && l0) ? mv * (lh * hw) : 0.f;
          se.o[3] = hhc * p.in_w + whc; se.w[3] = (t1 && l1) ? mv * (lh * lw) : 0.f;
        }
      }
      tab[e] = se;
    }
    __syncthreads();

    const int c_start = max(ci_lo, og * c_per_off), c_end = min(ci_hi, (og + 1) * c_per_off);
    const int k_start = (c_start - ci_lo) * KK, k_end = (c_end - ci_lo) * KK;   // within-group k range
    for (int k0 = k_start; k0 < k_end; k0 += BK) {
      // A slab: As[kk][m] = W[g*cout_g + oc0 + m][k0 + kk]
      for (int e = tid; e < BK * BM; e += DCN_THREADS) {
        const int m = e / BK, kk = e - m * BK;
        const int k = k0 + kk, oc = oc0 + m;
        float v = 0.f;
        if (k < k_end && oc < cout_g) v = to_acc(weight[((int64_t)(g * cout_g + oc)) * Kg + k]);
        As[kk * BMP + m] = v;
      }
      // B slab: Bs[kk][px] = sum_q w_q * in[ci][o_q]
      for (int e = tid; e < BK * BN; e += DCN_THREADS) {
        const int kk = e / BN, px = e - kk * BN;
        const int k = k0 + kk;
        float v = 0.f;
        if (k < k_end) {
          const int cil = k / KK, tap = k - cil * KK;
          const T* __restrict__ pl = in_b + (int64_t)(ci_lo + cil) * p.in_h * p.in_w;
          const SampleEnt se = tab[tap * BN + px];
          v = se.w[0] * to_acc(pl[se.o[0]]);
          v = fmaf(se.w[1], to_acc(pl[se.o[1]]), v);
          v = fmaf(se.w[2], to_acc(pl[se.o[2]]), v);
          v = fmaf(se.w[3], to_acc(pl[se.o[3]]), v);
        }
        Bs[kk * BN + px] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[8], bb[4];
        const float4 a0 = *reinterpret_cast<const float4*>(As + kk * BMP + tm * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(As + kk * BMP + tm * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + kk * BN + tn * 4);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // ---- epilogue: + bias, NCHW store ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int oc = oc0 + tm * 8 + i;
    if (oc >= cout_g) continue;
    const int ocg = g * cout_g + oc;
    const float bv = bias ? to_acc(bias[ocg]) : 0.f;
    T* __restrict__ o = out + ((int64_t)b * p.c_out + ocg) * HWo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pix = pix0 + tn * 4 + j;
      if (pix < HWo) o[pix] = from_acc<T, float>(acc[i][j] + bv);
    }
  }
}

template <typename T>
int launch_simt(const void* input, const void* weight, const void* offset, const void* mask, const void* bias, void* out,
                const DcnParams& p, cudaStream_t st) {
  const int KK = p.kh * p.kw;
  const size_t smem = (size_t)KK * BN * sizeof(SampleEnt) + (size_t)(BK * BMP + BK * BN) * 4;
  if (smem > (size_t)max_smem_optin() - 1024) {
    set_error("deform_conv2d: kernel %dx%d too large for the shared-memory sampling table", p.kh, p.kw);
    return VB200_EUNSUPPORTED;
  }
  if (smem > 48 * 1024)
    VB200_CUDA_TRY(ensure_dyn_smem<deform_conv2d_simt_kernel<T>>(smem));
  const int cout_g = p.c_out / p.groups;
  dim3 grid((unsigned)ceil_div(p.out_h * p.out_w, BN), (unsigned)(p.groups * ceil_div(cout_g, BM)), (unsigned)p.batch);
  deform_conv2d_simt_kernel<T><<<grid, DCN_THREADS, smem, st>>>((const T*)input, (const T*)weight, (const T*)offset,
                                                               (const T*)mask, (const T*)bias, (T*)out, p);
  return check_launch("deform_conv2d_simt_kernel");
}

// ---- float64 (the reference dispatches FLOATING_TYPES_AND_HALF; its gradcheck tests run in double) --------------------
// One thread per output element, double accumulation, bilinear_interpolate exactly as deform_conv2d_kernel.cu:97-134.
// A correctness path for tests, not a performance path.
__device__ __forceinline__ double dcn_bilinear_f64(const double* __restrict__ in, int H, int W, double h, double w) {
  if (h <= -1 || H <= h || w <= -1 || W <= w) return 0;
  const int hl = (int)floor(h), wl = (int)floor(w);
  const int hh_i = hl + 1, wh_i = wl + 1;
  const double lh = h - hl, lw = w - wl, hh = 1 - lh, hw = 1 - lw;
  const double v1 = (hl >= 0 && wl >= 0) ? in[hl * W + wl] : 0;
  const double v2 = (hl >= 0 && wh_i <= W - 1) ? in[hl * W + wh_i] : 0;
  const double v3 = (hh_i <= H - 1 && wl >= 0) ? in[hh_i * W + wl] : 0;
  const double v4 = (hh_i <= H - 1 && wh_i <= W - 1) ? in[hh_i * W + wh_i] : 0;
  return hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
}

__global__ void __launch_bounds__(256)
deform_conv2d_f64_kernel(const double* __restrict__ in, const double* __restrict__ w, const double* __restrict__ off,
                         const double* __restrict__ mask, const double* __restrict__ bias, double* __restrict__ out, DcnParams p) {
  const int HWo = p.out_h * p.out_w, HWi = p.in_h * p.in_w, KK = p.kh * p.kw;
  const int64_t total = (int64_t)p.batch * p.c_out * HWo;
  const int cin_g = p.c_in / p.groups, cout_g = p.c_out / p.groups, c_per_off = p.c_in / p.offset_groups;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int pix = (int)(idx % HWo);
    const int co = (int)((idx / HWo) % p.c_out);
    const int b = (int)(idx / HWo / p.c_out);
    const int oy = pix / p.out_w, ox = pix - oy * p.out_w;
    const int g = co / cout_g;
    double sum = 0;
    for (int ci = 0; ci < cin_g; ++ci) {
      const int c = g * cin_g + ci, og = c / c_per_off;
      const double* __restrict__ plane = in + ((int64_t)b * p.c_in + c) * HWi;
      for (int tap = 0; tap < KK; ++tap) {
        const int i = tap / p.kw, j = tap - i * p.kw;
        const int64_t ob = ((int64_t)b * p.offset_groups + og) * 2 * KK;
        const double y = (double)(oy * p.stride_h - p.pad_h + i * p.dil_h) + off[(ob + 2 * tap) * HWo + pix];
        const double x = (double)(ox * p.stride_w - p.pad_w + j * p.dil_w) + off[(ob + 2 * tap + 1) * HWo + pix];
        const double m = p.use_mask ? mask[(((int64_t)b * p.offset_groups + og) * KK + tap) * HWo + pix] : 1.0;
        sum += w[((int64_t)co * cin_g + ci) * KK + tap] * (m * dcn_bilinear_f64(plane, p.in_h, p.in_w, y, x));
      }
    }
    out[idx] = sum + (bias ? bias[co] : 0.0);
  }
}

}  // namespace

// deform_conv2d_tc.cu: returns 1 if handled, 0 if not applicable, other = error
int deform_conv2d_tc_try(const void* input, const void* weight, const void* offset, const void* mask, const void* bias,
                         void* out, int dtype, const DcnParams& p, void* workspace, size_t workspace_bytes, cudaStream_t st,
                         const DcnHints& hints);
size_t deform_conv2d_tc_workspace(int dtype, const DcnParams& p);
size_t deform_conv2d_tc_packed_bytes(int dtype, const DcnParams& p);
int deform_conv2d_tc_pack(const void* weight, void* packed, int dtype, const DcnParams& p, cudaStream_t st);

}  // namespace vb200

using namespace vb200;

static int dcn_out_dim(int in, int pad, int dil, int k, int stride) { return (in + 2 * pad - (dil * (k - 1) + 1)) / stride + 1; }

extern "C" size_t vb200_deform_conv2d_workspace_bytes(int dtype, int batch, int c_in, int in_h, int in_w, int c_out,
                                                      int kh, int kw, int out_h, int out_w, int groups,
                                                      int offset_groups) {
  DcnParams p{};
  p.batch = batch; p.c_in = c_in; p.in_h = in_h; p.in_w = in_w; p.c_out = c_out; p.kh = kh; p.kw = kw;
  p.groups = groups; p.offset_groups = offset_groups; p.out_h = out_h; p.out_w = out_w;
  return deform_conv2d_tc_workspace(dtype, p);
}

static int dcn_forward_impl(const void* input, const void* weight, const void* offset, const void* mask, const void* bias, void* out,
                            int dtype, int batch, int c_in, int in_h, int in_w, int c_out, int kh, int kw, int stride_h, int stride_w,
                            int pad_h, int pad_w, int dil_h, int dil_w, int groups, int offset_groups, int use_mask, void* workspace,
                            size_t workspace_bytes, vb200_stream stream, const DcnHints& hints);

extern "C" int vb200_deform_conv2d_forward(const void* input, const void* weight, const void* offset,
                                           const void* mask, const void* bias, void* out, int dtype, int batch,
                                           int c_in, int in_h, int in_w, int c_out, int kh, int kw, int stride_h,
                                           int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                           int offset_groups, int use_mask, void* workspace,
                                           size_t workspace_bytes, vb200_stream stream) {
  return dcn_forward_impl(input, weight, offset, mask, bias, out, dtype, batch, c_in, in_h, in_w, c_out, kh, kw, stride_h, stride_w, pad_h,
                          pad_w, dil_h, dil_w, groups, offset_groups, use_mask, workspace, workspace_bytes, stream, DcnHints{nullptr, 0, nullptr, 0, nullptr});
}

extern "C" size_t vb200_deform_conv2d_packed_weight_bytes(int dtype, int c_in, int c_out, int kh, int kw, int groups, int offset_groups) {
  DcnParams p{};
  p.batch = 1; p.c_in = c_in; p.in_h = 8; p.in_w = 8; p.c_out = c_out; p.kh = kh; p.kw = kw; p.groups = groups; p.offset_groups = offset_groups;
  return deform_conv2d_tc_packed_bytes(dtype, p);
}

extern "C" int vb200_deform_conv2d_pack_weight(const void* weight, void* packed, int dtype, int c_in, int c_out, int kh, int kw, int groups,
                                               int offset_groups, vb200_stream stream) {
  DcnParams p{};
  p.batch = 1; p.c_in = c_in; p.in_h = 8; p.in_w = 8; p.c_out = c_out; p.kh = kh; p.kw = kw; p.groups = groups; p.offset_groups = offset_groups;
  VB200_REQUIRE(weight && packed, "deform_conv2d_pack_weight: null pointer");
  VB200_REQUIRE(deform_conv2d_tc_packed_bytes(dtype, p) > 0, "deform_conv2d_pack_weight: this shape does not take the tensor-core path");
  return deform_conv2d_tc_pack(weight, packed, dtype, p, (cudaStream_t)stream);
}

extern "C" int vb200_deform_conv2d_forward_ex(const void* input, const void* weight, const void* packed_weight, int input_is_nhwc,
                                              const void* offset, const void* mask, const void* bias, void* out, int dtype, int batch,
                                              int c_in, int in_h, int in_w, int c_out, int kh, int kw, int stride_h, int stride_w, int pad_h,
                                              int pad_w, int dil_h, int dil_w, int groups, int offset_groups, int use_mask, void* workspace,
                                              size_t workspace_bytes, vb200_stream stream) {
  return dcn_forward_impl(input, weight, offset, mask, bias, out, dtype, batch, c_in, in_h, in_w, c_out, kh, kw, stride_h, stride_w, pad_h,
                          pad_w, dil_h, dil_w, groups, offset_groups, use_mask, workspace, workspace_bytes, stream,
                          DcnHints{packed_weight, input_is_nhwc, nullptr, 0, nullptr});
}

// deform_conv2d fused with the all-gather of its output: outs[0] is the caller's slot of its own gathered buffer, outs[1..n) the
// same slot of the peers' buffers (peer-mapped).  The tcgen05 kernel's epilogue stores each element to all of them; shapes that
// take another kernel are computed into outs[0] and copied to the peers on the same stream.
extern "C" int vb200_deform_conv2d_forward_gather(const void* input, const void* weight, const void* packed_weight, int input_is_nhwc,
                                                  const void* offset, const void* mask, const void* bias, void* const* outs, int n_outs,
                                                  int dtype, int batch, int c_in, int in_h, int in_w, int c_out, int kh, int kw,
                                                  int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int groups,
                                                  int offset_groups, int use_mask, void* workspace, size_t workspace_bytes,
                                                  vb200_stream stream) {
  VB200_REQUIRE(outs && n_outs >= 1 && n_outs <= 8, "deform_conv2d_gather: 1..8 destinations");
  for (int d = 0; d < n_outs; ++d) VB200_REQUIRE(outs[d] != nullptr, "deform_conv2d_gather: null destination");
  bool done = false;
  const int rc = dcn_forward_impl(input, weight, offset, mask, bias, outs[0], dtype, batch, c_in, in_h, in_w, c_out, kh, kw, stride_h,
                                  stride_w, pad_h, pad_w, dil_h, dil_w, groups, offset_groups, use_mask, workspace, workspace_bytes, stream,
                                  DcnHints{packed_weight, input_is_nhwc, outs + 1, n_outs - 1, &done});
  if (rc || done || n_outs == 1 || batch == 0 || c_out == 0) return rc;
  const size_t esize = dtype == VB200_F64 ? 8 : dtype == VB200_F32 ? 4 : 2;
  const size_t bytes = (size_t)batch * c_out * dcn_out_dim(in_h, pad_h, dil_h, kh, stride_h) * dcn_out_dim(in_w, pad_w, dil_w, kw, stride_w) * esize;
  for (int d = 1; d < n_outs; ++d) VB200_CUDA_TRY(cudaMemcpyAsync(outs[d], outs[0], bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

static int dcn_forward_impl(const void* input, const void* weight, const void* offset, const void* mask, const void* bias, void* out,
                            int dtype, int batch, int c_in, int in_h, int in_w, int c_out, int kh, int kw, int stride_h, int stride_w,
                            int pad_h, int pad_w, int dil_h, int dil_w, int groups, int offset_groups, int use_mask, void* workspace,
                            size_t workspace_bytes, vb200_stream stream, const DcnHints& hints) {
  // argument checks mirror deform_conv2d_kernel.cu:1056-1150
  VB200_REQUIRE(kh > 0 && kw > 0, "weight_h: %d weight_w: %d", kh, kw);
  VB200_REQUIRE(stride_h > 0 && stride_w > 0, "stride_h: %d stride_w: %d", stride_h, stride_w);
  VB200_REQUIRE(pad_h >= 0 && pad_w >= 0, "pad_h: %d pad_w: %d", pad_h, pad_w);
  VB200_REQUIRE(dil_h > 0 && dil_w > 0, "dilation_h: %d dilation_w: %d", dil_h, dil_w);
  VB200_REQUIRE(groups > 0 && offset_groups > 0 && c_in % groups == 0 && c_out % groups == 0 && c_in % offset_groups == 0,
                "deform_conv2d: channels not divisible by groups");
  DcnParams p{batch, c_in, in_h, in_w, c_out, kh, kw, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w,
              groups, offset_groups, use_mask, 0, 0};
  p.out_h = dcn_out_dim(in_h, pad_h, dil_h, kh, stride_h);
  p.out_w = dcn_out_dim(in_w, pad_w, dil_w, kw, stride_w);
  VB200_REQUIRE(p.out_h > 0 && p.out_w > 0, "Calculated output size too small - out_h: %d out_w: %d", p.out_h, p.out_w);
  if (batch == 0 || c_out == 0) return 0;
  VB200_REQUIRE(input && weight && offset && out && (!use_mask || mask), "deform_conv2d: null pointer");
  VB200_REQUIRE((int64_t)c_in * in_h * in_w < (1ll << 31) && batch <= 65535, "deform_conv2d: tensor too large");
  cudaStream_t st = (cudaStream_t)stream;
  const char* force = env_override(ENV_DCN_PATH);   // "simt" forces the SIMT kernel
  if (!(force && force[0] == 's')) {
    const int rc = deform_conv2d_tc_try(input, weight, offset, mask, bias, out, dtype, p, workspace, workspace_bytes, st, hints);
    if (rc != 0) return rc == 1 ? 0 : rc;
  }
  VB200_REQUIRE(!hints.input_is_nhwc, "deform_conv2d: a channels-last input is only accepted by the tensor-core path (this shape / dtype takes the SIMT kernel)");
  switch (dtype) {
    case VB200_F32: return launch_simt<float>(input, weight, offset, mask, bias, out, p, st);
    case VB200_F16: return launch_simt<__half>(input, weight, offset, mask, bias, out, p, st);
    case VB200_BF16: return launch_simt<__nv_bfloat16>(input, weight, offset, mask, bias, out, p, st);
    case VB200_F64: {
      const int64_t total = (int64_t)batch * c_out * p.out_h * p.out_w;
      const int grid = (int)(ceil_div64(total, 256) < (int64_t)sm_count() * 16 ? ceil_div64(total, 256) : (int64_t)sm_count() * 16);
      deform_conv2d_f64_kernel<<<grid, 256, 0, st>>>((const double*)input, (const double*)weight, (const double*)offset,
                                                    (const double*)mask, (const double*)bias, (double*)out, p);
      return check_launch("deform_conv2d_f64_kernel");
    }
  }
  set_error("deform_conv2d: unsupported dtype %d", dtype);
  return VB200_EUNSUPPORTED;
}
