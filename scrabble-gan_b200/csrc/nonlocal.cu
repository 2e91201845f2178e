// 1x1 projections of the non-local (self-attention) block, forward and backward, fused per data-flow step.
// Reference: arch_ops.py:38-46,55-57 (theta / phi / g projections), :63-67 (output projection, sigma * o + x).
//
// All of these are skinny GEMMs over p = n*h*w pixel rows with K, N in {32, 48, 64}: HBM-bound (each operand should
// move once), far too narrow for a 128-wide tensor-core tile, so they are FFMA kernels that stream a tile of rows
// through shared memory:
//   k_nl_rowgemm<K,N>  C[p,N] (+)= alpha * A[p,K] . W[K,N]      block = 128 rows x 2 column halves, W resident in smem,
//                                                                A tile staged with coalesced float4 loads
//   k_nl_wgrad<KA,KB>  dW[KA,KB] += alpha * A[p,KA]^T . B[p,KB]  reduction over rows: each block owns a slab of rows,
//                                                                keeps a 4x4 register tile per thread and adds its
//                                                                partial dW with one atomicAdd per element
// A, B, C and W may be "column-segmented": the logical matrix is the concatenation of up to three separate row-major
// arrays (theta | phi | g), which is how the three projections are computed from ONE pass over x.
#include "common.cuh"

struct SegMat {
  float* p[3];
  int w[3];      // widths (multiples of 4); unused segments have w = 0
};

__device__ __forceinline__ void seg_find(const SegMat& m, int col, int& s, int& off) {
  s = 0;
  off = col;
  if (off >= m.w[0]) { off -= m.w[0]; s = 1; if (off >= m.w[1]) { off -= m.w[1]; s = 2; } }
}

#define NL_ROWS 128

// W logical [K,N]; if TRANS_W the stored matrix is the column-segmented [N, K] one and is read transposed.
template <int K, int N, bool TRANS_W>
__global__ void __launch_bounds__(256) k_nl_rowgemm(long long rows, SegMat A, SegMat W, SegMat C, const float* __restrict__ alpha_p,
                                                     int accumulate, const float* __restrict__ resid_scale_p,
                                                     const float* __restrict__ resid_x, float* __restrict__ resid_out) {
  __shared__ float As[NL_ROWS][K + 1];
  __shared__ __align__(16) float Ws[K][N];
  const int tid = threadIdx.x;
  // ---- weights -> smem (once per block) ----
  for (int i = tid; i < K * N; i += 256) {
    int k = i / N, n = i % N;
    int s, off;
    float v;
    if (!TRANS_W) {            // stored [K, N] column-segmented over n
      seg_find(W, n, s, off);
      v = W.p[s][(long long)k * W.w[s] + off];
    } else {                   // stored [N, K] column-segmented over k
      seg_find(W, k, s, off);
      v = W.p[s][(long long)n * W.w[s] + off];
    }
    Ws[k][n] = v;
  }
  const float alpha = alpha_p ? *alpha_p : 1.f;
  const float rscale = resid_scale_p ? *resid_scale_p : 1.f;
  const int r = tid & (NL_ROWS - 1), half = tid >> 7;
  constexpr int NH = N / 2;

  for (long long row0 = (long long)blockIdx.x * NL_ROWS; row0 < rows; row0 += (long long)gridDim.x * NL_ROWS) {
    const int cnt = rows - row0 < NL_ROWS ? (int)(rows - row0) : NL_ROWS;
    __syncthreads();
    // ---- A tile -> smem: every segment's tile is one contiguous chunk of cnt * w floats ----
    int koff = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int w = A.w[s];
      if (w == 0) continue;
      const float4* src = reinterpret_cast<const float4*>(A.p[s] + row0 * w);
      const int n4 = cnt * w / 4;
      for (int i = tid; i < n4; i += 256) {
        float4 v = src[i];
        int e = i * 4, rr = e / w, cc = e % w + koff;
        As[rr][cc] = v.x; As[rr][cc + 1] = v.y; As[rr][cc + 2] = v.z; As[rr][cc + 3] = v.w;
      }
      koff += w;
    }
    __syncthreads();
    if (r < cnt) {
      float acc[NH];
#pragma unroll
      for (int j = 0; j < NH; ++j) acc[j] = 0.f;
#pragma unroll 8
      for (int k = 0; k < K; ++k) {
        const float a = As[r][k];
        const float4* wp = reinterpret_cast<const float4*>(&Ws[k][half * NH]);
#pragma unroll
        for (int j = 0; j < NH / 4; ++j) {
          float4 w4 = wp[j];
          acc[4 * j] = fmaf(a, w4.x, acc[4 * j]);
          acc[4 * j + 1] = fmaf(a, w4.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(a, w4.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(a, w4.w, acc[4 * j + 3]);
        }
      }
      const long long row = row0 + r;
#pragma unroll
      for (int j = 0; j < NH; j += 4) {
        int col = half * NH + j, s, off;
        seg_find(C, col, s, off);
        float4 v = make_float4(alpha * acc[j], alpha * acc[j + 1], alpha * acc[j + 2], alpha * acc[j + 3]);
        float* cp = C.p[s] + row * C.w[s] + off;
        if (accumulate) {
          float4 o = sg_ld4(cp);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        sg_st4(cp, v);
        if (resid_out) {       // second output: resid_out = rscale * C + x   (sigma * og + x, arch_ops.py:67)
          float4 x = sg_ld4(resid_x + row * N + col);
          sg_st4(resid_out + row * N + col, make_float4(fmaf(rscale, v.x, x.x), fmaf(rscale, v.y, x.y), fmaf(rscale, v.z, x.z),
                                                        fmaf(rscale, v.w, x.w)));
        }
      }
    }
  }
}

#define NLW_ROWS 64
template <int KA, int KB>
__global__ void __launch_bounds__((KA / 4) * (KB / 4)) k_nl_wgrad(long long rows, long long rows_per_block, SegMat A, SegMat B,
                                                                   SegMat DW, const float* __restrict__ alpha_p) {
  constexpr int TA = KA / 4, TB = KB / 4, NT = TA * TB;
  __shared__ __align__(16) float As[NLW_ROWS][KA];
  __shared__ __align__(16) float Bs[NLW_ROWS][KB];
  const int tid = threadIdx.x;
  const int ti = tid % TA, tj = tid / TA;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  for (long long row0 = r0; row0 < r1; row0 += NLW_ROWS) {
    const int cnt = r1 - row0 < NLW_ROWS ? (int)(r1 - row0) : NLW_ROWS;
    __syncthreads();
    int koff = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int w = A.w[s];
      if (w == 0) continue;
      const float4* src = reinterpret_cast<const float4*>(A.p[s] + row0 * w);
      for (int i = tid; i < cnt * w / 4; i += NT) {
        int e = i * 4, rr = e / w, cc = e % w + koff;
        *reinterpret_cast<float4*>(&As[rr][cc]) = src[i];
      }
      koff += w;
    }
    koff = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int w = B.w[s];
      if (w == 0) continue;
      const float4* src = reinterpret_cast<const float4*>(B.p[s] + row0 * w);
      for (int i = tid; i < cnt * w / 4; i += NT) {
        int e = i * 4, rr = e / w, cc = e % w + koff;
        *reinterpret_cast<float4*>(&Bs[rr][cc]) = src[i];
      }
      koff += w;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < cnt; ++rr) {
      float4 a = *reinterpret_cast<const float4*>(&As[rr][4 * ti]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[rr][4 * tj]);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]); acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]); acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
      acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]); acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
      acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]); acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
    }
  }
  const float alpha = alpha_p ? *alpha_p : 1.f;
  int s, off;
  seg_find(DW, 4 * tj, s, off);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* dp = DW.p[s] + (long long)(4 * ti + i) * DW.w[s] + off;
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dp + j, alpha * acc[i][j]);
  }
}

static SegMat seg1(const float* a, int w) {
  SegMat m;
  m.p[0] = const_cast<float*>(a); m.p[1] = m.p[2] = nullptr;
  m.w[0] = w; m.w[1] = m.w[2] = 0;
  return m;
}
static SegMat seg3(const float* a, int wa, const float* b, int wb, const float* c, int wc) {
  SegMat m;
  m.p[0] = const_cast<float*>(a); m.p[1] = const_cast<float*>(b); m.p[2] = const_cast<float*>(c);
  m.w[0] = wa; m.w[1] = wb; m.w[2] = wc;
  return m;
}
static int nl_grid(sg_ctx* ctx, long long rows) {
  long long need = (rows + NL_ROWS - 1) / NL_ROWS, cap = (long long)ctx->num_sms * 4;
  return (int)(need < cap ? need : cap);
}
static void nl_wgrad_grid(sg_ctx* ctx, long long rows, int* grid, long long* rpb) {
  long long blocks = (long long)ctx->num_sms * 2;
  long long r = (rows + blocks - 1) / blocks;
  r = (r + NLW_ROWS - 1) / NLW_ROWS * NLW_ROWS;
  *rpb = r;
  *grid = (int)((rows + r - 1) / r);
}

#define NL_C 64
#define NL_DK 8
#define NL_DV 32
#define NL_ALIGNED(p) ((((uintptr_t)(p)) & 15) == 0)

extern "C" {

int sg_nonlocal_proj_fwd(sg_ctx* ctx, const float* x, long long rows, const float* w_theta, const float* w_phi,
                         const float* w_g, float* theta, float* phi_f, float* g_f) {
  SG_REQUIRE(ctx && x && w_theta && w_phi && w_g && theta && phi_f && g_f, "sg_nonlocal_proj_fwd: NULL");
  SG_REQUIRE(NL_ALIGNED(x) && NL_ALIGNED(theta) && NL_ALIGNED(phi_f) && NL_ALIGNED(g_f), "sg_nonlocal_proj_fwd: 16-byte alignment");
  if (rows == 0) return SG_OK;
  k_nl_rowgemm<NL_C, 2 * NL_DK + NL_DV, false><<<nl_grid(ctx, rows), 256, 0, ctx->stream>>>(
      rows, seg1(x, NL_C), seg3(w_theta, NL_DK, w_phi, NL_DK, w_g, NL_DV), seg3(theta, NL_DK, phi_f, NL_DK, g_f, NL_DV), nullptr, 0,
      nullptr, nullptr, nullptr);
  SG_POST_LAUNCH(ctx);
  return SG_OK;
}

int sg_nonlocal_out_fwd(sg_ctx* ctx, const float* o, long long rows, const float* w_o, const float* sigma, const float* x,
                        float* og, float* out) {
  SG_REQUIRE(ctx && o && w_o && sigma && x && og && out, "sg_nonlocal_out_fwd: NULL");
  SG_REQUIRE(NL_ALIGNED(o) && NL_ALIGNED(x) && NL_ALIGNED(og) && NL_ALIGNED(out), "sg_nonlocal_out_fwd: 16-byte alignment");
  if (rows == 0) return SG_OK;
  // og = o . Wo (kept un-scaled for d sigma = <dout, og>);  out = sigma * og + x
  k_nl_rowgemm<NL_DV, NL_C, false><<<nl_grid(ctx, rows), 256, 0, ctx->stream>>>(rows, seg1(o, NL_DV), seg1(w_o, NL_C), seg1(og, NL_C),
                                                                                 nullptr, 0, sigma, x, out);
  SG_POST_LAUNCH(ctx);
  return SG_OK;
}

int sg_nonlocal_out_bwd(sg_ctx* ctx, const float* dout, const float* o, long long rows, const float* w_o, const float* sigma,
                        float* d_o, float* dw_o) {
  SG_REQUIRE(ctx && dout && o && w_o && sigma && d_o, "sg_nonlocal_out_bwd: NULL");
  SG_REQUIRE(NL_ALIGNED(dout) && NL_ALIGNED(o) && NL_ALIGNED(d_o), "sg_nonlocal_out_bwd: 16-byte alignment");
  if (rows == 0) return SG_OK;
  // d_o = sigma * dout . Wo^T
  k_nl_rowgemm<NL_C, NL_DV, true><<<nl_grid(ctx, rows), 256, 0, ctx->stream>>>(rows, seg1(dout, NL_C), seg1(w_o, NL_C), seg1(d_o, NL_DV),
                                                                                sigma, 0, nullptr, nullptr, nullptr);
  SG_POST_LAUNCH(ctx);
  if (dw_o) {                  // dWo[32,64] += sigma * o^T . dout
    int grid;
    long long rpb;
    nl_wgrad_grid(ctx, rows, &grid, &rpb);
    k_nl_wgrad<NL_DV, NL_C><<<grid, (NL_DV / 4) * (NL_C / 4), 0, ctx->stream>>>(rows, rpb, seg1(o, NL_DV), seg1(dout, NL_C),
                                                                                 seg1(dw_o, NL_C), sigma);
    SG_POST_LAUNCH(ctx);
  }
  return SG_OK;
}

int sg_nonlocal_proj_bwd(sg_ctx* ctx, const float* x, const float* dtheta, const float* dphi_f, const float* dg_f, long long rows,
                         const float* w_theta, const float* w_phi, const float* w_g, float* dx, float* dw_theta, float* dw_phi,
                         float* dw_g) {
  SG_REQUIRE(ctx && x && dtheta && dphi_f && dg_f && w_theta && w_phi && w_g && dx, "sg_nonlocal_proj_bwd: NULL");
  SG_REQUIRE(NL_ALIGNED(x) && NL_ALIGNED(dtheta) && NL_ALIGNED(dphi_f) && NL_ALIGNED(dg_f) && NL_ALIGNED(dx),
             "sg_nonlocal_proj_bwd: 16-byte alignment");
  SG_REQUIRE((dw_theta != nullptr) == (dw_phi != nullptr) && (dw_phi != nullptr) == (dw_g != nullptr),
             "sg_nonlocal_proj_bwd: give all three filter gradients or none");
  if (rows == 0) return SG_OK;
  SegMat d = seg3(dtheta, NL_DK, dphi_f, NL_DK, dg_f, NL_DV);
  // dx += [dtheta | dphi | dg] . [Wtheta | Wphi | Wg]^T
  k_nl_rowgemm<2 * NL_DK + NL_DV, NL_C, true><<<nl_grid(ctx, rows), 256, 0, ctx->stream>>>(
      rows, d, seg3(w_theta, NL_DK, w_phi, NL_DK, w_g, NL_DV), seg1(dx, NL_C), nullptr, 1, nullptr, nullptr, nullptr);
  SG_POST_LAUNCH(ctx);
  if (dw_theta) {              // [dWtheta | dWphi | dWg] += x^T . [dtheta | dphi | dg]
    int grid;
    long long rpb;
    nl_wgrad_grid(ctx, rows, &grid, &rpb);
    k_nl_wgrad<NL_C, 2 * NL_DK + NL_DV><<<grid, (NL_C / 4) * ((2 * NL_DK + NL_DV) / 4), 0, ctx->stream>>>(
        rows, rpb, seg1(x, NL_C), d, seg3(dw_theta, NL_DK, dw_phi, NL_DK, dw_g, NL_DV), nullptr);
    SG_POST_LAUNCH(ctx);
  }
  return SG_OK;
}

}  // extern "C"
