// ctx.cu — context, error reporting, kernel profiler, device-wide scan, device vectors.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

static thread_local std::string g_thread_err;

void ff_set_thread_error(const std::string &s) { g_thread_err = s; }

void ff_report_error(ffcuda_ctx *ctx, const char *msg)
{
    g_thread_err = msg ? msg : "unknown error";
    if (ctx) ctx->err = g_thread_err;
}

static thread_local ffcuda_ctx *g_cur_ctx = nullptr;
ffcuda_ctx *ff_current_ctx() { return g_cur_ctx; }
void ff_enter(ffcuda_ctx *ctx)
{
    g_cur_ctx = ctx;
    if (ctx) cudaSetDevice(ctx->device);
}

int ff_nloc(int dim, int order) { return order == 1 ? dim + 1 : (dim == 2 ? 6 : 10); }

void ff_prof_flush(ffcuda_ctx *ctx)
{
    if (ctx->prof_pending.empty()) return;
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &pp : ctx->prof_pending) {
        float ms = 0;
        FF_CUDA(cudaEventElapsedTime(&ms, pp.e0, pp.e1));
        ProfEntry &e = ctx->prof_acc[pp.name];
        e.ms += ms;
        e.count++;
        cudaEventDestroy(pp.e0);
        cudaEventDestroy(pp.e1);
    }
    ctx->prof_pending.clear();
}

extern "C" int ffcuda_ctx_create(int device, ffcuda_ctx **out)
{
    ffcuda_ctx *ctx = nullptr;
    FF_API_BEGIN
    FF_REQUIRE(out, "ffcuda_ctx_create: null output");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw FFError(std::string("no CUDA device available (ffcuda has no CPU fallback): ") +
                      cudaGetErrorString(e));
    FF_REQUIRE(device >= 0 && device < ndev, "device index out of range");
    FF_CUDA(cudaSetDevice(device));
    ctx = new ffcuda_ctx();
    ctx->device = device;
    try {
        FF_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
        ctx->stream = ctx->own_stream;
        cudaDeviceProp prop;
        FF_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        if (const char *e = getenv("FFCUDA_TILES")) ctx->tile_policy = std::max(0, std::min(2, atoi(e)));
        if (const char *e = getenv("FFCUDA_TILE_FANS")) ctx->tile_fans = atoi(e) != 0;
        if (const char *e = getenv("FFCUDA_TILE_ROWS")) ctx->tile_rows = std::max(8, std::min(256, atoi(e)));
        FF_CUDA(cudaMalloc((void **)&ctx->d_scal, 256 * sizeof(double))); // [0,64): CG scalars and flags, [64, ..): P2PDesc
        // zeroed on the context's own (non-blocking) stream: the legacy stream does not order with it
        FF_CUDA(cudaMemsetAsync(ctx->d_scal, 0, 256 * sizeof(double), ctx->stream));
        FF_CUDA(cudaStreamSynchronize(ctx->stream));
        FF_CUDA(cudaMallocHost((void **)&ctx->h_scal, 64 * sizeof(double)));
    } catch (...) { // nothing of a half-built context survives an error
        if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
        if (ctx->d_scal) cudaFree(ctx->d_scal);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        ctx = nullptr;
        cudaGetLastError();
        throw;
    }
    *out = ctx;
    FF_API_END(ctx)
}

// ---------------------------------------------------------------------------------------------------
// caching device allocator
// ---------------------------------------------------------------------------------------------------
static size_t pool_round(size_t bytes) { return bytes < (1u << 20) ? (bytes + 511) & ~(size_t)511 : (bytes + ((1u << 20) - 1)) & ~(size_t)((1u << 20) - 1); }

void ff_pool_trim(ffcuda_ctx *ctx)
{
    if (ctx->pool_free.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->pool_free) cudaFree(kv.second);
    ctx->pool_total -= ctx->pool_cached;
    ctx->pool_cached = 0;
    ctx->pool_free.clear();
}

void *ff_pool_alloc(ffcuda_ctx *ctx, size_t bytes)
{
    const size_t want = pool_round(bytes);
    // smallest cached block that fits without wasting more than a quarter of it
    auto it = ctx->pool_free.lower_bound(want);
    if (it != ctx->pool_free.end() && it->first <= want + want / 4 + (1u << 20)) {
        void *p = it->second;
        ctx->pool_cached -= it->first;
        ctx->pool_live[p] = it->first;
        ctx->pool_free.erase(it);
        return p;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { // give the cache back to the driver and retry once
        cudaGetLastError();
        ff_pool_trim(ctx);
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess)
        throw FFError(std::string("out of device memory allocating ") + std::to_string(want) + " bytes: " + cudaGetErrorString(e));
    ctx->pool_live[p] = want;
    ctx->pool_total += want;
    return p;
}

void ff_pool_free(ffcuda_ctx *ctx, void *p)
{
    auto it = ctx->pool_live.find(p);
    if (it == ctx->pool_live.end()) return;
    const size_t sz = it->second;
    ctx->pool_live.erase(it);
    ctx->pool_free.emplace(sz, p);
    ctx->pool_cached += sz;
}

static void ctx_teardown(ffcuda_ctx *ctx)
{
    cudaSetDevice(ctx->device);
    try { ff_prof_flush(ctx); } catch (...) {}
    ff_comm_release(ctx);
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->pool_free) cudaFree(kv.second);
    for (auto &kv : ctx->pool_live) cudaFree(kv.first);
    if (ctx->d_scal) cudaFree(ctx->d_scal);
    if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
    if (ctx->d_partial) cudaFree(ctx->d_partial);
    if (ctx->copy_stream) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        cudaEventDestroy(ctx->copy_event);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (g_cur_ctx == ctx) g_cur_ctx = nullptr;
    delete ctx;
}

void ff_ctx_unref(ffcuda_ctx *ctx)
{
    if (--ctx->refs <= 0 && ctx->closed) ctx_teardown(ctx);
}

extern "C" void ffcuda_ctx_destroy(ffcuda_ctx *ctx)
{
    if (!ctx || ctx->closed) return;
    ctx->closed = true;
    if (ctx->refs <= 0) ctx_teardown(ctx); // otherwise the last handle still alive tears it down
}

extern "C" const char *ffcuda_last_error(ffcuda_ctx *ctx)
{
    if (ctx) return ctx->err.c_str();
    return g_thread_err.c_str();
}

extern "C" int ffcuda_ctx_sync(ffcuda_ctx *ctx)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx, "null context");
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) FF_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    FF_API_END(ctx)
}

extern "C" int ffcuda_ctx_set_stream(ffcuda_ctx *ctx, void *cuda_stream)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx, "null context");
    FF_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    FF_API_END(ctx)
}

extern "C" void *ffcuda_ctx_get_stream(ffcuda_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int ffcuda_ctx_set_option(ffcuda_ctx *ctx, const char *name, int value)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && name, "ffcuda_ctx_set_option: bad arguments");
    const std::string n(name);
    if (n == "tile_policy") {
        FF_REQUIRE(value >= 0 && value <= 2, "tile_policy must be 0, 1 or 2");
        ctx->tile_policy = value;
    } else if (n == "tile_fans") {
        FF_REQUIRE(value == 0 || value == 1, "tile_fans must be 0 or 1");
        ctx->tile_fans = value;
    } else if (n == "tile_rows") {
        FF_REQUIRE(value >= 8 && value <= 256, "tile_rows must be in 8..256");
        ctx->tile_rows = value;
    } else if (n == "gmres_coop") {
        FF_REQUIRE(value == 0 || value == 1, "gmres_coop must be 0 or 1");
        ctx->gmres_coop = value;
    } else
        throw FFError("ffcuda_ctx_set_option: unknown option '" + n + "'");
    FF_API_END(ctx)
}

extern "C" int ffcuda_prof_enable(ffcuda_ctx *ctx, int on)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx, "null context");
    ff_prof_flush(ctx);
    ctx->prof = on != 0;
    FF_API_END(ctx)
}

extern "C" int ffcuda_prof_reset(ffcuda_ctx *ctx)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx, "null context");
    ff_prof_flush(ctx);
    ctx->prof_acc.clear();
    ctx->launches = 0;
    FF_API_END(ctx)
}

extern "C" int ffcuda_prof_get(ffcuda_ctx *ctx, const char *prefix, double *ms, int64_t *launches)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx, "null context");
    ff_prof_flush(ctx);
    double t = 0;
    int64_t c = 0;
    size_t lp = prefix ? strlen(prefix) : 0;
    for (auto &kv : ctx->prof_acc)
        if (lp == 0 || kv.first.compare(0, lp, prefix) == 0) {
            t += kv.second.ms;
            c += kv.second.count;
        }
    if (ms) *ms = t;
    if (launches) *launches = c;
    FF_API_END(ctx)
}

extern "C" int64_t ffcuda_launch_count(ffcuda_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------------
// Device-wide exclusive scan of int32 (three passes: per-tile sums, scan of tile sums, per-tile scan).
// Totals are accumulated in 64 bits so that an nnz overflow of int32 is detected, not wrapped.
// ---------------------------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 8;                        // per thread
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS; // 2048

__device__ __forceinline__ long long warp_incl_scan(long long v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ long long block_excl_scan(long long v, long long *total)
{
    __shared__ long long wsum[SCAN_THREADS / 32];
    __shared__ long long tot;
    long long inc = warp_incl_scan(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        long long s = l < SCAN_THREADS / 32 ? wsum[l] : 0;
        long long si = warp_incl_scan(s);
        if (l < SCAN_THREADS / 32) wsum[l] = si - s;
        if (l == SCAN_THREADS / 32 - 1) tot = si;
    }
    __syncthreads();
    long long r = inc - v + wsum[w];
    *total = tot;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const int32_t *__restrict__ in, size_t n, long long *__restrict__ tsum)
{
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    long long tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) tsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_offsets(long long *tsum, int ntiles, long long *total)
{
    // single block: sequential over chunks of SCAN_THREADS tiles
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b = 0; b < ntiles; b += SCAN_THREADS) {
        int i = b + threadIdx.x;
        long long v = i < ntiles ? tsum[i] : 0;
        long long tot;
        long long ex = block_excl_scan(v, &tot);
        if (i < ntiles) tsum[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int32_t *__restrict__ in, int32_t *__restrict__ out, size_t n,
                                                              const long long *__restrict__ toff)
{
    // each thread owns SCAN_ITEMS consecutive items (blocked arrangement) so the output order is the input order
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    long long tot;
    long long ex = block_excl_scan(s, &tot) + toff[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = (int32_t)ex;
        ex += v[i];
    }
}

// Single-pass variant (decoupled look-back): tiles take tickets from an atomic counter, publish (flag, sum) words —
// flag 1: the tile's own sum, flag 2: inclusive prefix — and a tile resolves its offset by walking back over its
// predecessors' words.  One launch and one read of the input instead of three launches and two reads.
// st[0..ntiles): status words, st[ntiles]: grand total, st[ntiles+1]: ticket counter; all zero on entry.
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_lookback(const int32_t *__restrict__ in, int32_t *__restrict__ out, size_t n,
                                                                 int ntiles, unsigned long long *__restrict__ st)
{
    __shared__ int s_tile;
    __shared__ long long s_prefix;
    constexpr unsigned long long MASK = (1ull << 62) - 1ull;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(st + ntiles + 1, 1ull);
    __syncthreads();
    const int tile = s_tile;
    const size_t base = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    long long tot;
    long long ex = block_excl_scan(s, &tot);
    if (threadIdx.x == 0 && tile > 0) atomicExch(st + tile, (1ull << 62) | (unsigned long long)tot);
    if (threadIdx.x < 32) { // warp 0 walks back over the predecessors' words, 32 at a time
        const int lane = threadIdx.x;
        long long prefix = 0;
        for (int j = tile - 1; tile > 0;) {
            const int idx = j - lane;
            unsigned long long w = 2ull << 62; // before the first tile: inclusive prefix 0
            if (idx >= 0) {
                do {
                    w = *reinterpret_cast<volatile unsigned long long *>(st + idx);
                } while ((w >> 62) == 0);
            }
            const unsigned incl = __ballot_sync(0xffffffffu, (w >> 62) == 2);
            const int first = incl ? __ffs(incl) - 1 : 31;
            long long val = lane <= first ? (long long)(w & MASK) : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            prefix += val;
            if (incl) break;
            j -= 32;
        }
        if (lane == 0) {
            __threadfence();
            atomicExch(st + tile, (2ull << 62) | (unsigned long long)(prefix + tot));
            if (tile == ntiles - 1) st[ntiles] = (unsigned long long)(prefix + tot);
            s_prefix = prefix;
        }
    }
    __syncthreads();
    ex += s_prefix;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = (int32_t)ex;
        ex += v[i];
    }
}

void ff_exclusive_scan_i32(ffcuda_ctx *ctx, const int32_t *in, int32_t *out, size_t n, int64_t *total)
{
    if (n > 0 && in != out) { // (the three-pass version below also works in place)
        const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
        DBuf<unsigned long long> stw;
        stw.alloc((size_t)ntiles + 2);
        cudaStream_t st = ctx->stream;
        FF_CUDA(cudaMemsetAsync(stw.p, 0, stw.bytes(), st));
        ff_launch(ctx, "scan_lookback", [&] { k_scan_lookback<<<ntiles, SCAN_THREADS, 0, st>>>(in, out, n, ntiles, stw.p); });
        unsigned long long tot = 0;
        FF_CUDA(cudaMemcpyAsync(&tot, stw.p + ntiles, sizeof(tot), cudaMemcpyDeviceToHost, st));
        FF_CUDA(cudaStreamSynchronize(st));
        if (total) *total = (int64_t)tot;
        return;
    }
    if (n == 0) {
        if (total) *total = 0;
        return;
    }
    int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    DBuf<long long> tsum;
    tsum.alloc((size_t)ntiles + 1);
    cudaStream_t st = ctx->stream;
    ff_launch(ctx, "scan_tile_sums", [&] { k_scan_tile_sums<<<ntiles, SCAN_THREADS, 0, st>>>(in, n, tsum.p); });
    ff_launch(ctx, "scan_tile_offsets", [&] { k_scan_tile_offsets<<<1, SCAN_THREADS, 0, st>>>(tsum.p, ntiles, tsum.p + ntiles); });
    ff_launch(ctx, "scan_tiles", [&] { k_scan_tiles<<<ntiles, SCAN_THREADS, 0, st>>>(in, out, n, tsum.p); });
    long long tot = 0;
    FF_CUDA(cudaMemcpyAsync(&tot, tsum.p + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, st));
    FF_CUDA(cudaStreamSynchronize(st));
    if (total) *total = tot;
}

// ---------------------------------------------------------------------------------------------------
// device vectors
// ---------------------------------------------------------------------------------------------------
__global__ void k_fill(double *p, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

extern "C" int ffcuda_vec_create(ffcuda_ctx *ctx, int n, ffcuda_vec **out)
{
    FF_API_BEGIN
    FF_REQUIRE(ctx && out && n >= 0, "ffcuda_vec_create: bad arguments");
    ff_enter(ctx);
    ffcuda_vec *v = new ffcuda_vec();
    v->ctx = ctx;
    v->ref.set(ctx);
    v->n = n;
    v->d.alloc((size_t)n);
    if (n) FF_CUDA(cudaMemsetAsync(v->d.p, 0, v->d.bytes(), ctx->stream));
    *out = v;
    FF_API_END(ctx)
}

extern "C" int ffcuda_vec_upload(ffcuda_vec *v, const double *host)
{
    FF_API_BEGIN
    FF_REQUIRE(v && host, "ffcuda_vec_upload: null argument");
    FF_CUDA(cudaMemcpyAsync(v->d.p, host, v->d.bytes(), cudaMemcpyHostToDevice, v->ctx->stream));
    FF_CUDA(cudaStreamSynchronize(v->ctx->stream));
    FF_API_END(v ? v->ctx : nullptr)
}

extern "C" int ffcuda_vec_download(ffcuda_vec *v, double *host)
{
    FF_API_BEGIN
    FF_REQUIRE(v && host, "ffcuda_vec_download: null argument");
    FF_CUDA(cudaMemcpyAsync(host, v->d.p, v->d.bytes(), cudaMemcpyDeviceToHost, v->ctx->stream));
    FF_CUDA(cudaStreamSynchronize(v->ctx->stream));
    FF_API_END(v ? v->ctx : nullptr)
}

extern "C" int ffcuda_vec_fill(ffcuda_vec *v, double value)
{
    FF_API_BEGIN
    FF_REQUIRE(v, "ffcuda_vec_fill: null argument");
    ffcuda_ctx *ctx = v->ctx;
    if (v->n) {
        int blocks = min(ff_blocks((size_t)v->n, 256), ctx->sm_count * 8);
        ff_launch(ctx, "vec_fill", [&] { k_fill<<<blocks, 256, 0, ctx->stream>>>(v->d.p, (size_t)v->n, value); });
    }
    FF_API_END(v ? v->ctx : nullptr)
}

extern "C" void *ffcuda_vec_ptr(ffcuda_vec *v) { return v ? (void *)v->d.p : nullptr; }

extern "C" void ffcuda_vec_destroy(ffcuda_vec *v)
{
    if (!v) return;
    ff_enter(v->ctx);
    delete v;
}
