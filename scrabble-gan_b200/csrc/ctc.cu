// Fused softmax -> log(p + 1e-7) -> CTC loss -> gradient w.r.t. the Dense pre-activations (K15).
// Reference: net_architecture.py:55 (Dense softmax) and :57-72 (K.ctc_batch_cost):
//     y_pred = log(softmax(z) + 1e-7);  loss = tf.nn.ctc_loss(inputs=y_pred, ...)   (which re-softmaxes its inputs)
// => effective per-frame distribution q = (p + eps) / (1 + C eps), blank = C-1, repeats merged.
// One CTA per sample; thread s owns extended-label state s (S = 2L+1); log-space alpha/beta in shared memory.
// dL/du (u = log(p+eps)) = q - occupancy/P  [tf.nn.ctc_loss gradient], chained through log and softmax to z.
#include "common.cuh"

#define CTC_EPS 1e-7f
#define NEG_INF (-INFINITY)

// accurate (not fast-intrinsic) log-sum-exp: the recursion runs T steps in the log domain and its output is
// differenced against log q when forming posteriors, so absolute log-domain error matters (CTC parity 1e-4)
__device__ __forceinline__ float lse2(float a, float b) {
  float m = fmaxf(a, b);
  if (m == NEG_INF) return NEG_INF;
  return m + log1pf(expf(fminf(a, b) - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  float m = fmaxf(fmaxf(a, b), c);
  if (m == NEG_INF) return NEG_INF;
  // sum of the two non-max terms, then log1p
  float s = expf(a - m) + expf(b - m) + expf(c - m) - 1.f;
  return m + log1pf(s);
}

// dynamic smem: logq[T*C] | alpha[T*S] | beta[T*S] | occ[C] | ext[S] (int)
// Ragged batches (K.ctc_batch_cost's own interface: per-sample input_length / label_length, net_architecture.py:57-72):
// input_len / label_len, when given, hold every sample's frame count T_b <= Tmax and label count L_b <= Lmax; logits / labels
// / grad keep the rectangular [b, Tmax, C] / [b, Lmax] layout, frames t >= T_b get a zero gradient.
__global__ void k_ctc(const float* __restrict__ logits, const int* __restrict__ labels, int Tmax, int C, int Lmax,
                      const int* __restrict__ input_len, const int* __restrict__ label_len, float* __restrict__ loss,
                      float* __restrict__ grad) {
  sg_pdl_prologue();
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  int T = input_len ? input_len[b] : Tmax, L = label_len ? label_len[b] : Lmax;
  T = T < 1 ? 1 : (T > Tmax ? Tmax : T);
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  const int S = 2 * L + 1, Smax = 2 * Lmax + 1;
  float* logq = sm;
  float* alpha = logq + Tmax * C;
  float* beta = alpha + Tmax * Smax;
  float* occ = beta + Tmax * Smax;
  int* ext = reinterpret_cast<int*>(occ + C);
  __shared__ float s_logp;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int blank = C - 1;
  const float* z = logits + (long long)b * Tmax * C;

  for (int s = tid; s < S; s += blockDim.x) ext[s] = (s & 1) ? labels[b * Lmax + (s >> 1)] : blank;

  // per-frame effective log-probabilities: one warp per frame
  const float log_norm = logf(1.f + (float)C * CTC_EPS);
  for (int t = warp; t < T; t += nwarps) {
    float mx = NEG_INF;
    for (int k = lane; k < C; k += 32) mx = fmaxf(mx, z[t * C + k]);
    mx = sg_warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < C; k += 32) sum += expf(z[t * C + k] - mx);
    sum = sg_warp_sum(sum);
    float inv = 1.f / sum;
    for (int k = lane; k < C; k += 32) {
      float p = expf(z[t * C + k] - mx) * inv;
      logq[t * C + k] = logf(p + CTC_EPS) - log_norm;
    }
  }
  __syncthreads();

  // alpha recursion
  for (int s = tid; s < S; s += blockDim.x)
    alpha[s] = (s < 2) ? logq[ext[s]] : NEG_INF;
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    for (int s = tid; s < S; s += blockDim.x) {
      const float* prev = alpha + (t - 1) * S;
      float a0 = prev[s];
      float a1 = s >= 1 ? prev[s - 1] : NEG_INF;
      float a2 = (s >= 2 && ext[s] != blank && ext[s] != ext[s - 2]) ? prev[s - 2] : NEG_INF;
      float v = lse3(a0, a1, a2);
      alpha[t * S + s] = (v == NEG_INF) ? NEG_INF : v + logq[t * C + ext[s]];
    }
    __syncthreads();
  }
  // beta recursion (beta_t(s) includes the emission at t)
  for (int s = tid; s < S; s += blockDim.x)
    beta[(T - 1) * S + s] = (s >= S - 2) ? logq[(T - 1) * C + ext[s]] : NEG_INF;
  __syncthreads();
  for (int t = T - 2; t >= 0; --t) {
    for (int s = tid; s < S; s += blockDim.x) {
      const float* nxt = beta + (t + 1) * S;
      float b0 = nxt[s];
      float b1 = s + 1 < S ? nxt[s + 1] : NEG_INF;
      float b2 = (s + 2 < S && ext[s + 2] != blank && ext[s + 2] != ext[s]) ? nxt[s + 2] : NEG_INF;
      float v = lse3(b0, b1, b2);
      beta[t * S + s] = (v == NEG_INF) ? NEG_INF : v + logq[t * C + ext[s]];
    }
    __syncthreads();
  }
  if (tid == 0) {
    float a = alpha[(T - 1) * S + S - 1];
    float c = S >= 2 ? alpha[(T - 1) * S + S - 2] : NEG_INF;
    float lp = lse2(a, c);
    s_logp = lp;
    loss[b] = -lp;
  }
  __syncthreads();
  if (!grad) return;
  const float logp = s_logp;
  float* gz = grad + (long long)b * Tmax * C;
  for (int i = T * C + tid; i < Tmax * C; i += blockDim.x) gz[i] = 0.f;      // frames beyond this sample's length
  if (logp == NEG_INF) {                                                    // no valid alignment (T_b too short): loss = +inf
    for (int i = tid; i < T * C; i += blockDim.x) gz[i] = 0.f;
    return;
  }

  // gradient, frame by frame (block-cooperative: occupancy scatter, then a block reduction for the softmax chain)
  __shared__ float red[32];
  for (int t = 0; t < T; ++t) {
    // occupancy of class k = sum over the extended-label positions s that carry k, in position order (no atomics: the
    // blank and repeated letters occur at several positions, and a fixed order keeps the gradient bitwise repeatable)
    for (int k = tid; k < C; k += blockDim.x) {
      float o = 0.f;
      for (int s = 0; s < S; ++s) {
        if (ext[s] != k) continue;
        float ab = alpha[t * S + s] + beta[t * S + s];
        if (ab != NEG_INF) o += expf(ab - 2.f * logq[t * C + k] - logp);   // alpha*beta/q^2/P  (beta includes one q)
      }
      occ[k] = o;
    }
    __syncthreads();
    // g_u[k] = q - q*occ' where occ' = sum alpha*beta/(q^2 P) ... so occupancy/P expressed relative to q:
    //   dL/du_k = q_k - (1/P) sum_s alpha_t(s) beta_t(s) / q_k = q_k - q_k * occ[k]
    float partial = 0.f;
    for (int k = tid; k < C; k += blockDim.x) {
      float q = expf(logq[t * C + k]);
      float gu = q - q * occ[k];
      float p = q * (1.f + (float)C * CTC_EPS) - CTC_EPS;
      if (p < 0.f) p = 0.f;
      float gp = gu / (p + CTC_EPS);
      occ[k] = gp;                      // reuse as dL/dp
      partial += p * gp;
    }
    float dot = sg_block_sum(partial, red);
    for (int k = tid; k < C; k += blockDim.x) {
      float q = expf(logq[t * C + k]);
      float p = q * (1.f + (float)C * CTC_EPS) - CTC_EPS;
      if (p < 0.f) p = 0.f;
      gz[t * C + k] = p * (occ[k] - dot);
    }
    __syncthreads();
  }
}

static int ctc_launch(sg_ctx* ctx, const float* logits, const int* labels, int b, int t, int c, int l, const int* input_len,
                      const int* label_len, float* loss, float* grad_logits, const char* who) {
  SG_REQUIRE(ctx && logits && labels && loss, "%s: NULL", who);
  SG_REQUIRE(b >= 0 && t > 0 && c > 1 && l >= 0, "%s: bad sizes", who);
  SG_REQUIRE(input_len || label_len || t >= l, "%s: %d frames cannot emit %d labels", who, t, l);
  if (b == 0) return SG_OK;
  int S = 2 * l + 1;
  size_t smem = sizeof(float) * ((size_t)t * c + 2 * (size_t)t * S + c) + sizeof(int) * S;
  SG_REQUIRE(smem <= 200 * 1024, "%s: T*C too large for shared memory (%zu bytes)", who, smem);
  if (smem > 48 * 1024) SG_CHECK_CUDA(cudaFuncSetAttribute(k_ctc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sg_launch(ctx, k_ctc, b, 128, smem, logits, labels, t, c, l, input_len, label_len, loss, grad_logits);
  SG_POST_LAUNCH(ctx);
  return SG_OK;
}

extern "C" int sg_ctc(sg_ctx* ctx, const float* logits, const int* labels, int b, int t, int c, int l, float* loss,
                      float* grad_logits) {
  return ctc_launch(ctx, logits, labels, b, t, c, l, nullptr, nullptr, loss, grad_logits, "sg_ctc");
}

/* ragged batch: per-sample frame counts input_len[b] <= t_max and label counts label_len[b] <= l_max (device int32) */
extern "C" int sg_ctc_ragged(sg_ctx* ctx, const float* logits, const int* labels, int b, int t_max, int c, int l_max,
                             const int* input_len, const int* label_len, float* loss, float* grad_logits) {
  SG_REQUIRE(input_len && label_len, "sg_ctc_ragged: NULL length arrays");
  return ctc_launch(ctx, logits, labels, b, t_max, c, l_max, input_len, label_len, loss, grad_logits, "sg_ctc_ragged");
}
