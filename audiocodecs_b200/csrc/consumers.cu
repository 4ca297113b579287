// Token consumers: the two HBM-bound kernels that sit right after the tokenizer in the reference's downstream recipes
// (SURVEY 8f-4).
//   ac_token_histogram    : per-codebook code counts of a [rows][K] int64 token tensor -- the accumulation step of
//                           CodebookUtil.append (R/downstream/metrics/codebook_util.py:39-49: K x `unique(return_counts)` +
//                           host-side index-add there; one pass over the tokens here, 8 bytes per token).
//   ac_multihead_embedding: out[row][k][:] = weight[toks[row][k] + offset[k]][:]  (R/downstream/models/multihead.py:28-69:
//                           `input + offsets` -> F.embedding; `vocab_size` as a token value selects the padding row).
#include "common.cuh"

namespace {

// block-private histogram of one codebook column in shared memory, merged with one atomic per non-empty bin
__global__ void token_histogram_kernel(const int64_t* __restrict__ toks, long long rows, int K, int vocab, long long* __restrict__ counts,
                                       int* __restrict__ err_flag) {
    extern __shared__ unsigned int hist[];  // [vocab]
    const int k = blockIdx.y;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
        const long long t = __ldg(toks + r * K + k);
        if (t >= 0 && t < vocab) atomicAdd(&hist[(int)t], 1u);
        else if (err_flag) atomicExch(err_flag, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < vocab; i += blockDim.x)
        if (hist[i]) atomicAdd(reinterpret_cast<unsigned long long*>(counts + (long long)k * vocab + i), (unsigned long long)hist[i]);
}

// one thread per 4 consecutive output floats
__global__ void multihead_embedding_kernel(const int64_t* __restrict__ toks, const float* __restrict__ weight,
                                           const int64_t* __restrict__ offsets, float* __restrict__ out, long long rows, int K, int D,
                                           long long vocab, long long padding_row, long long n_embeddings, int* __restrict__ err_flag) {
    const int per = D / 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long item = gid / per;  // (row, k)
    if (item >= rows * K) return;
    const int k = (int)(item % K);
    const int d0 = (int)(gid % per) * 4;
    const long long t = __ldg(toks + item);
    long long idx = t + __ldg(offsets + k);
    if (padding_row >= 0 && t == vocab) idx = padding_row;  // the padding token of every head maps to the shared padding row
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && idx >= 0 && idx < n_embeddings) v = __ldg(reinterpret_cast<const float4*>(weight + idx * D + d0));
    else if (err_flag) atomicExch(err_flag, 1);
    *reinterpret_cast<float4*>(out + item * D + d0) = v;
}

}  // namespace

extern "C" int ac_token_histogram(const int64_t* toks, int64_t rows, int32_t num_codebooks, int32_t vocab_size, int64_t* counts,
                                  int32_t* err_flag, void* stream) {
    AC_REQUIRE(toks && counts && rows > 0 && num_codebooks > 0 && num_codebooks <= 65535, "ac_token_histogram: bad arguments");
    AC_REQUIRE(vocab_size > 0 && vocab_size <= 48 * 1024 / 4, "ac_token_histogram: vocab_size %d (at most 12288 bins of shared memory)", vocab_size);
    long long blocks = (rows + 256 * 64 - 1) / (256 * 64);  // ~64 tokens per thread: few merges
    if (blocks < 1) blocks = 1;
    if (blocks > 296) blocks = 296;
    token_histogram_kernel<<<dim3((unsigned)blocks, num_codebooks), 256, (size_t)vocab_size * 4, (cudaStream_t)stream>>>(
        toks, rows, num_codebooks, vocab_size, (long long*)counts, err_flag);
    return ac::finish_launch("ac_token_histogram");
}

extern "C" int ac_multihead_embedding(const int64_t* toks, const float* weight, const int64_t* offsets, float* out, int64_t rows,
                                      int32_t num_codebooks, int32_t dim, int64_t vocab_size, int64_t padding_row,
                                      int64_t num_embeddings, int32_t* err_flag, void* stream) {
    AC_REQUIRE(toks && weight && offsets && out && rows > 0 && num_codebooks > 0, "ac_multihead_embedding: bad arguments");
    AC_REQUIRE(dim > 0 && dim % 4 == 0 && ((uintptr_t)weight & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "ac_multihead_embedding: dim %d must be a multiple of 4 and the tensors 16-byte aligned", dim);
    const long long total = rows * num_codebooks * (dim / 4);
    multihead_embedding_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        toks, weight, offsets, out, rows, num_codebooks, dim, vocab_size, padding_row, num_embeddings, err_flag);
    return ac::finish_launch("ac_multihead_embedding");
}
