// HOST-resident operands: batched multiply + relinearize (+ rescale) whose inputs and result live in
// host memory, the C-ABI form of ExecutionOptions::set_storage_type(storage_type::HOST) for the hot
// path (reference: src/include/heongpu/util/storagemanager.cuh:113-167, truth table README.md:349-366;
// the reference moves each operand with its own blocking copy, one ciphertext at a time).
//
// The batch is cut into chunks; three streams run  H2D(chunk k+1) | compute(chunk k) | D2H(chunk k-1)
// over two sets of device staging buffers, so the link is busy in both directions while the SMs work
// and the call is bounded by max(PCIe time, compute time) instead of their sum.  The call is
// asynchronous with respect to the host: it is ordered after `stream` at entry and `stream` is ordered
// after it at exit.  Calls issued on different caller streams overlap (the pipeline's three internal streams are
// shared per host thread, so their chunks simply queue behind each other).
#include "ops.hpp"

namespace heon {

namespace {
struct Pipe {
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr}, ev_entry = nullptr, ev_exit = nullptr;
    int device = -1;
    void init(int dev)
    {
        if (device == dev)
            return;
        device = dev;
        cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&s_cmp, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
        for (int i = 0; i < 2; ++i)
        {
            cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev_cmp[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&ev_entry, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_exit, cudaEventDisableTiming);
    }
};
thread_local Pipe t_pipe; // one pipeline per host thread (the reference's threading contract: one stream per thread)
} // namespace

// h_a, h_b: [batch][2][L][N]; h_out: [batch][2][Lout][N] with Lout = L - (rescale ? 1 : 0)
void op_mulrelin_host(const Context& c, const u64* h_a, const u64* h_b, u64* h_out, const u64* relin_key, int depth,
                      int rescale, int batch, int chunk, cudaStream_t st)
{
    const int L = c.Q_size - depth;
    if (depth < 0 || L < 1)
        throw std::invalid_argument("invalid depth");
    if (rescale && L < 2)
        throw std::logic_error("Ciphertext modulus can not be dropped!");
    if (chunk < 1)
        chunk = 4;
    chunk = std::min(chunk, batch);
    const long long N = c.n;
    const int Lout = L - (rescale ? 1 : 0);
    const size_t in_words = (size_t) 2 * L * N, ct3_words = (size_t) 3 * L * N, out_words = (size_t) 2 * Lout * N;
    Pipe& p = t_pipe;
    p.init(c.device);
    cudaEventRecord(p.ev_entry, st);
    for (cudaStream_t s : {p.s_in, p.s_cmp, p.s_out})
        cudaStreamWaitEvent(s, p.ev_entry, 0);
    // two sets of staging buffers: inputs a, b and the 3-component product (the result is its first two components)
    u64* stage = nullptr;
    const size_t set_words = (size_t) chunk * (2 * in_words + ct3_words);
    if (cudaMallocAsync(&stage, 2 * set_words * 8, p.s_in) != cudaSuccess)
        throw std::runtime_error("cudaMallocAsync failed (host pipeline staging)");
    cudaEventRecord(p.ev_in[0], p.s_in); // the allocation is visible to the other streams through this event
    cudaStreamWaitEvent(p.s_cmp, p.ev_in[0], 0);
    cudaStreamWaitEvent(p.s_out, p.ev_in[0], 0);
    const int n_chunks = (batch + chunk - 1) / chunk;
    bool used[2] = {false, false};
    for (int k = 0; k < n_chunks; ++k)
    {
        const int j = k & 1;
        const int b0 = k * chunk, cb = std::min(chunk, batch - b0);
        u64* dA = stage + j * set_words;
        u64* dB = dA + (size_t) chunk * in_words;
        u64* dC = dB + (size_t) chunk * in_words;
        if (used[j])
            cudaStreamWaitEvent(p.s_in, p.ev_cmp[j], 0); // chunk k-2 has consumed this input set
        cudaMemcpyAsync(dA, h_a + (size_t) b0 * in_words, (size_t) cb * in_words * 8, cudaMemcpyHostToDevice, p.s_in);
        cudaMemcpyAsync(dB, h_b + (size_t) b0 * in_words, (size_t) cb * in_words * 8, cudaMemcpyHostToDevice, p.s_in);
        cudaEventRecord(p.ev_in[j], p.s_in);
        cudaStreamWaitEvent(p.s_cmp, p.ev_in[j], 0);
        if (used[j])
            cudaStreamWaitEvent(p.s_cmp, p.ev_out[j], 0); // chunk k-2's result has left this product buffer
        op_multiply(c, dA, (long long) in_words, dB, (long long) in_words, dC, (long long) ct3_words, depth, cb, p.s_cmp);
        op_relinearize(c, dC, (long long) ct3_words, relin_key, depth, cb, p.s_cmp);
        if (rescale)
            op_rescale(c, dC, (long long) ct3_words, depth, cb, p.s_cmp);
        cudaEventRecord(p.ev_cmp[j], p.s_cmp);
        cudaStreamWaitEvent(p.s_out, p.ev_cmp[j], 0);
        cudaMemcpy2DAsync(h_out + (size_t) b0 * out_words, out_words * 8, dC, ct3_words * 8, out_words * 8, cb,
                          cudaMemcpyDeviceToHost, p.s_out);
        cudaEventRecord(p.ev_out[j], p.s_out);
        used[j] = true;
    }
    // The D2H stream finishes last (its copy of the last chunk follows that chunk's compute, which follows its H2D,
    // and the earlier chunks precede them in stream order): free the staging there and order the caller's stream
    // after it.  Nothing is queued behind the pipeline on the H2D or compute streams, so a following call on ANOTHER
    // caller stream (the reference's multi-stream usage, example/basic/9_multi_stream_usage_way1.cpp) starts its
    // copies while this call's last chunks are still computing / leaving: no fill / drain bubble between calls.
    cudaFreeAsync(stage, p.s_out);
    cudaEventRecord(p.ev_exit, p.s_out);
    cudaStreamWaitEvent(st, p.ev_exit, 0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw std::runtime_error(std::string("host pipeline: ") + cudaGetErrorString(e));
}

} // namespace heon
