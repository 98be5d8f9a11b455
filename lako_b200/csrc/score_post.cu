// Device-side per-fact segmentation of the step-0 cross-attention token scores: the post-processing the reference
// runs as a Python loop per example after a (B, N, L) device->host copy (/root/reference/src/model.py:166-203:
// find_index :100-105, get_attention_score :107-115) plus the per-fact softmax of test_reader.py:75-76.
//
// Semantics kept exactly (they are integer / ordering logic, so segment boundaries and counts are bit-exact):
//   * scores come from the FIRST passage of the second half of the passages (torch.chunk(.., 2, dim=1)[1][:, 0]), token
//     ids from passage 1 (context_ids[b][1]);
//   * a fact is the token range [start, end) with start = 2 for the first fact ("fact:" is two tokens) and
//     end = 1 + index of the next id 5 ('.'); at most n_context facts; if fewer were found and the passage's last id is
//     not pad (0), the unterminated tail [start, L) is one more fact;
//   * mean | max | "21mean" (mean of the ceil(len/2) largest); missing facts are -5; everything / divisor;
//   * arithmetic in float64 in the reference's order (Python floats: sequential left-to-right sums; heapq.nlargest
//     sums its result in descending order), so the values agree with the host loop to the last bit given the same
//     token scores.
#include "ops.cuh"

namespace lako {

namespace {

constexpr int FS_MAX_L = 2048;

struct FactParams {
  const float* tok_a;      // [B][N*L] fp32
  const float* tok_b;      // optional second summand (all layers = first half + last half), fp32 add as torch does
  const long long* ids1;   // [B][ld_ids] token ids of passage 1
  double* out;             // [B][n_context]
  int L, ld_ids, score_off, tok_ld, n_context, style, softmax;
  double divisor;
};

__global__ void __launch_bounds__(256) fact_scores_kernel(const FactParams p) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double s_sc[FS_MAX_L];
  __shared__ int s_id5[FS_MAX_L];     // 1 where the token id is 5
  __shared__ int s_beg[FS_MAX_L / 2 + 2], s_end[FS_MAX_L / 2 + 2];
  __shared__ int s_nseg;
  __shared__ double s_val[FS_MAX_L / 2 + 2];
  const int b = blockIdx.x, tid = threadIdx.x, L = p.L;
  const long long* ids = p.ids1 + static_cast<size_t>(b) * p.ld_ids;
  const float* ta = p.tok_a + static_cast<size_t>(b) * p.tok_ld + p.score_off;
  const float* tb = p.tok_b ? p.tok_b + static_cast<size_t>(b) * p.tok_ld + p.score_off : nullptr;
  for (int i = tid; i < L; i += blockDim.x) {
    float v = ta[i];
    if (tb) v = v + tb[i];
    s_sc[i] = static_cast<double>(v);
    s_id5[i] = ids[i] == 5 ? 1 : 0;
  }
  __syncthreads();
  if (tid == 0) {
    // model.py:178-191 -- boundaries, sequentially as the reference finds them
    int n = 0, start = 2;
    while (n < p.n_context) {
      int j = start;
      while (j < L && !s_id5[j]) ++j;  // fact_list.index(5, start)
      if (j >= L) break;
      s_beg[n] = start;
      s_end[n] = j + 1;
      ++n;
      start = j + 1;
    }
    if (n < p.n_context && ids[L - 1] != 0 && L > start) {
      s_beg[n] = start;
      s_end[n] = L;
      ++n;
    }
    s_nseg = n;
  }
  __syncthreads();
  const int nseg = s_nseg;
  for (int f = tid; f < nseg; f += blockDim.x) {
    const int s0 = s_beg[f], s1 = s_end[f], len = s1 - s0;
    double v;
    if (p.style == 0) {  // mean
      double acc = 0.0;
      for (int i = s0; i < s1; ++i) acc += s_sc[i];
      v = acc / static_cast<double>(len);
    } else if (p.style == 1) {  // max
      v = s_sc[s0];
      for (int i = s0 + 1; i < s1; ++i) v = fmax(v, s_sc[i]);
    } else {  // 21mean: mean of the max(int((len + 1) / 2), 1) largest, summed in descending order
      for (int i = s0 + 1; i < s1; ++i) {  // insertion sort, descending, inside this fact's own (disjoint) range
        const double x = s_sc[i];
        int j = i - 1;
        while (j >= s0 && s_sc[j] < x) {
          s_sc[j + 1] = s_sc[j];
          --j;
        }
        s_sc[j + 1] = x;
      }
      int num = (len + 1) / 2;
      if (num < 1) num = 1;
      double acc = 0.0;
      for (int i = 0; i < num; ++i) acc += s_sc[s0 + i];
      v = acc / static_cast<double>(num);
    }
    s_val[f] = v / p.divisor;
  }
  __syncthreads();
  double* out = p.out + static_cast<size_t>(b) * p.n_context;
  if (!p.softmax) {
    for (int f = tid; f < p.n_context; f += blockDim.x) out[f] = f < nseg ? s_val[f] : -5.0 / p.divisor;
    return;
  }
  // test_reader.py:76 torch.softmax(scores, dim=-1) over all n_context entries (the -5 fillers included), float64
  if (tid == 0) {
    const double filler = -5.0 / p.divisor;
    double m = -INFINITY;
    for (int f = 0; f < p.n_context; ++f) m = fmax(m, f < nseg ? s_val[f] : filler);
    double z = 0.0;
    for (int f = 0; f < p.n_context; ++f) z += exp((f < nseg ? s_val[f] : filler) - m);
    for (int f = 0; f < p.n_context; ++f) out[f] = exp((f < nseg ? s_val[f] : filler) - m) / z;
  }
}

}  // namespace

// style: 0 mean, 1 max, 2 "21mean".  score_passage: index of the passage whose token scores are segmented.
int fact_scores(const float* tok_a, const float* tok_b, const long long* ids1, int ld_ids, double* out, int B, int N, int L,
                int score_passage, int n_context, int style, double divisor, int softmax, cudaStream_t s) {
  LAKO_REQUIRE(tok_a && ids1 && out && B > 0 && N > 0 && L >= 1 && L <= FS_MAX_L, LAKO_ERR_UNSUPPORTED);
  LAKO_REQUIRE(n_context >= 1 && n_context <= FS_MAX_L / 2 && style >= 0 && style <= 2 && divisor != 0.0, LAKO_ERR_BAD_ARG);
  LAKO_REQUIRE(score_passage >= 0 && score_passage < N, LAKO_ERR_BAD_ARG);
  FactParams p{tok_a, tok_b, ids1, out, L, ld_ids, score_passage * L, N * L, n_context, style, softmax, divisor};
  launch_pdl(fact_scores_kernel, dim3(B), dim3(256), 0, s, p);
  LAKO_CHECK_LAUNCH();
  return LAKO_OK;
}

}  // namespace lako

extern "C" int lako_fact_scores(const float* tok_a, const float* tok_b, const int64_t* ids1, int ld_ids, double* out, int B,
                                int N, int L, int score_passage, int n_context, int style, double divisor, int softmax,
                                void* stream) {
  int rc = lako::check_arch();
  if (rc) return rc;
  return lako::fact_scores(tok_a, tok_b, reinterpret_cast<const long long*>(ids1), ld_ids, out, B, N, L, score_passage,
                           n_context, style, divisor, softmax, static_cast<cudaStream_t>(stream));
}
