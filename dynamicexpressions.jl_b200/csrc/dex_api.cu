// dex_api.cu — the C ABI of libdexb200.so (include/dexb200.h): contexts, operator
// tables, packed populations and the evaluation entry points.  Everything above this
// file (Julia shim, Python mirror) sees only `extern "C"` functions with raw pointers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/dexb200.h"
#include "dex_kernels.h"
// NVTX ranges around the entry points (the reference has no tracing of its own; a profiler that is
// attached — ncu --nvtx, Nsight Systems — sees `dexb200::<entry point>`; without one a range costs a
// few nanoseconds).  Header-only: nvtx3 loads the tool's injection library lazily, nothing to link.
#include <nvtx3/nvToolsExt.h>
namespace {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace
#define DEX_RANGE(name) NvtxRange nvtx_range_("dexb200::" name)
#include "dex_tape.h"

using namespace dex;

// ---- opaque types ------------------------------------------------------------------
struct dex_optable {
    OpTable t;
};

struct dex_ctx {
    int device = -1;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int sm_count = 148;
    std::string last_error;
    // scratch owned by the context
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    void* xt = nullptr;      // feature-major padded copy of X for the interpreter
    size_t xt_bytes = 0;
    void* dev_io = nullptr;  // device staging for the *_host entry points
    size_t dev_io_bytes = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // recorded behind the device work of every compute call: the context's scratch buffers (xt,
    // scratch, dev_io) are shared by all calls, so a call issued on ANOTHER stream must first wait
    // for the previous one (dex_ctx_set_stream)
    cudaEvent_t ev_done = nullptr;
    bool ev_done_valid = false;
    int64_t launches = 0;    // kernels launched through this context
};

struct dex_population {
    PackedPopulation h;  // host image
    int device = -1;
    Instr* d_tape = nullptr;
    int64_t* d_tape_off = nullptr;
    int32_t* d_const_ord = nullptr;
    int64_t* d_const_off = nullptr;
    int64_t* d_const_pos = nullptr;
    // folded image (h.folded): what dex_eval* run
    Instr* d_ftape = nullptr;
    int64_t* d_ftape_off = nullptr;
    int64_t* d_fconst_pos = nullptr;
    Instr* d_ctape = nullptr;
    int64_t* d_seg = nullptr;
    int64_t* d_seg_off = nullptr;
    // outcome of the constant folding for the CURRENT constants, per rule (0 = evaluation,
    // 1 = gradient): computed by launch_fold on first use, invalidated when constants change
    // all device arrays above live in ONE allocation (eleven cudaMalloc/cudaFree pairs per packed
    // population cost more than flattening it): d_blob owns, the typed pointers are views
    void* d_blob = nullptr;
    size_t blob_cap = 0;
    uint8_t* d_fold_ok[2] = {nullptr, nullptr};   // views into d_blob
    bool fold_valid[2] = {false, false};
    // tree-chunk tables (n_chunks + 1 tree indices, balanced by tape length): two slots inside
    // d_blob, keyed by n_chunks; a third key overwrites the older slot (stream-ordered upload)
    int32_t* d_chunk_slot[2] = {nullptr, nullptr};
    int32_t chunk_key[2] = {-1, -1};
    int chunk_next = 0;
    std::map<int32_t, std::vector<int32_t>> chunk_tables_host;
};

namespace {

// ---- device block pool -------------------------------------------------------------------------
// cudaMalloc / cudaFree cost 0.1 - 5 ms each (mapping and unmapping physical memory, an implicit
// device synchronisation) — more than flattening a 1 000-tree population.  Callers whose trees
// change every generation pack and destroy populations continuously, so population storage comes
// from a per-device free list of blocks that is never returned to the driver below a cap.
struct DevPool {
    std::mutex m;
    std::multimap<size_t, void*> free_blocks;   // capacity -> block
    size_t cached = 0;
};
constexpr int kMaxDevices = 64;
constexpr size_t kPoolCap = (size_t)1 << 30;    // bytes kept per device
DevPool g_pool[kMaxDevices];

size_t pool_round(size_t bytes) {   // 64 KiB granules up to 1 MiB, then 1 MiB granules
    const size_t g = bytes <= ((size_t)1 << 20) ? ((size_t)64 << 10) : ((size_t)1 << 20);
    return ((std::max<size_t>(bytes, 1) + g - 1) / g) * g;
}
cudaError_t pool_alloc(int dev, size_t bytes, void** out, size_t* cap) {
    *cap = pool_round(bytes);
    if (dev >= 0 && dev < kMaxDevices) {
        DevPool& p = g_pool[dev];
        std::lock_guard<std::mutex> lk(p.m);
        auto it = p.free_blocks.lower_bound(*cap);
        if (it != p.free_blocks.end() && it->first <= 2 * *cap + ((size_t)1 << 20)) {
            *out = it->second;
            *cap = it->first;
            p.cached -= it->first;
            p.free_blocks.erase(it);
            return cudaSuccess;
        }
    }
    return cudaMalloc(out, *cap);
}
// the caller guarantees that no device work still uses the block
void pool_free(int dev, void* ptr, size_t cap) {
    if (!ptr) return;
    if (dev >= 0 && dev < kMaxDevices) {
        DevPool& p = g_pool[dev];
        std::lock_guard<std::mutex> lk(p.m);
        if (p.cached + cap <= kPoolCap) {
            p.free_blocks.emplace(cap, ptr);
            p.cached += cap;
            return;
        }
    }
    cudaFree(ptr);
}

int set_err(dex_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}
int cuda_err(dex_ctx* ctx, cudaError_t e, const char* what) {
    return set_err(ctx, DEX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, call)                                              \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_err(ctx, e__, #call);  \
    } while (0)

struct OpInfo { int code; int degree; const char* sym; const char* name; const char* aliases; };
const OpInfo kOps[] = {
#define DEX_OP(SYM, code, degree, name, aliases) {code, degree, #SYM, name, aliases},
#include "../../include/dex_ops.def"
#undef DEX_OP
};

const OpInfo* find_op(int code) {
    for (const OpInfo& o : kOps)
        if (o.code == code) return &o;
    return nullptr;
}

bool alias_match(const char* aliases, const char* name) {
    const size_t n = std::strlen(name);
    const char* p = aliases;
    while (*p) {
        const char* q = std::strchr(p, '|');
        size_t len = q ? (size_t)(q - p) : std::strlen(p);
        if (len == n && std::strncmp(p, name, n) == 0) return true;
        if (!q) break;
        p = q + 1;
    }
    return false;
}

int ensure_device(dex_ctx* ctx) {
    if (!ctx) return DEX_ERR_INVALID;
    if (ctx->device < 0) return set_err(ctx, DEX_ERR_CUDA, "host-only context: no CUDA device bound (libdexb200 has no CPU fallback)");
    CU(ctx, cudaSetDevice(ctx->device));
    return DEX_OK;
}

// marks the end of the device work of a compute call on the current stream (see dex_ctx::ev_done)
int finish_call(dex_ctx* ctx, int rc) {
    if (rc == DEX_OK && ctx->ev_done) {
        CU(ctx, cudaEventRecord(ctx->ev_done, ctx->stream));
        ctx->ev_done_valid = true;
    }
    return rc;
}
// work enqueued on `next` from now on runs after everything this context has enqueued so far
int order_after_previous(dex_ctx* ctx, cudaStream_t next) {
    if (ctx->device < 0 || next == ctx->stream || !ctx->ev_done_valid) return DEX_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamWaitEvent(next, ctx->ev_done, 0));
    return DEX_OK;
}

int ensure_scratch(dex_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return DEX_OK;
    if (ctx->scratch) { CU(ctx, cudaStreamSynchronize(ctx->stream)); CU(ctx, cudaFree(ctx->scratch)); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
    CU(ctx, cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return DEX_OK;
}
int ensure_xt(dex_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->xt_bytes) return DEX_OK;
    if (ctx->xt) { CU(ctx, cudaStreamSynchronize(ctx->stream)); CU(ctx, cudaFree(ctx->xt)); ctx->xt = nullptr; ctx->xt_bytes = 0; }
    CU(ctx, cudaMalloc(&ctx->xt, bytes));
    ctx->xt_bytes = bytes;
    return DEX_OK;
}
int ensure_dev_io(dex_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->dev_io_bytes) return DEX_OK;
    if (ctx->dev_io) { CU(ctx, cudaStreamSynchronize(ctx->stream)); CU(ctx, cudaFree(ctx->dev_io)); ctx->dev_io = nullptr; ctx->dev_io_bytes = 0; }
    CU(ctx, cudaMalloc(&ctx->dev_io, bytes));
    ctx->dev_io_bytes = bytes;
    return DEX_OK;
}
int ensure_pinned(dex_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes) return DEX_OK;
    if (ctx->pinned) { CU(ctx, cudaStreamSynchronize(ctx->stream)); CU(ctx, cudaFreeHost(ctx->pinned)); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
    CU(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return DEX_OK;
}

template <typename U>
int upload(dex_ctx* ctx, U** dptr, const std::vector<U>& v) {
    *dptr = nullptr;
    // 64 elements of slack: the interpreter prefetches tape lines a little past the end
    const size_t bytes = (std::max<size_t>(v.size(), 1) + 64) * sizeof(U);
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(dptr), bytes));
    if (!v.empty()) CU(ctx, cudaMemcpyAsync(*dptr, v.data(), v.size() * sizeof(U), cudaMemcpyHostToDevice, ctx->stream));
    return DEX_OK;
}

// Every device array of a packed population in one allocation.  Each array keeps the 64 elements
// of slack behind its last entry that the interpreters rely on (tape line prefetch, offsets read
// one tree ahead) and starts on a 256-byte boundary.
int upload_population(dex_ctx* ctx, dex_population* pop) {
    const PackedPopulation& h = pop->h;
    const PackedPopulation& f = *h.folded;
    struct Item { void** dst; const void* src; size_t bytes, padded, off; };
    std::vector<Item> items;
    size_t total = 0;
    auto add_raw = [&](void** dptr, const void* src, size_t bytes, size_t elems, size_t esz) {
        Item it;
        it.dst = dptr;
        it.src = src;
        it.bytes = bytes;
        it.padded = (((std::max<size_t>(elems, 1) + 64) * esz) + 255) & ~(size_t)255;
        it.off = total;
        total += it.padded;
        items.push_back(it);
    };
    auto add = [&](auto** dptr, const auto& vec) {
        using U = typename std::remove_reference<decltype(vec)>::type::value_type;
        add_raw(reinterpret_cast<void**>(dptr), vec.data(), vec.size() * sizeof(U), vec.size(), sizeof(U));
    };
    add(&pop->d_tape, h.tape);
    add(&pop->d_tape_off, h.tape_off);
    add(&pop->d_const_ord, h.tape_const_ord);
    add(&pop->d_const_off, h.const_off);
    add(&pop->d_const_pos, h.const_pos);
    add(&pop->d_ftape, f.tape);
    add(&pop->d_ftape_off, f.tape_off);
    add(&pop->d_fconst_pos, f.const_pos);
    add(&pop->d_ctape, f.ctape);
    add(&pop->d_seg, f.seg);
    add(&pop->d_seg_off, f.seg_off);
    const size_t uploaded = total;     // everything above is copied from the host image
    const size_t nt = (size_t)h.n_trees;
    for (int k = 0; k < 2; ++k) add_raw(reinterpret_cast<void**>(&pop->d_fold_ok[k]), nullptr, 0, nt, 1);
    for (int k = 0; k < 2; ++k) add_raw(reinterpret_cast<void**>(&pop->d_chunk_slot[k]), nullptr, 0, nt + 1, sizeof(int32_t));
    cudaError_t e = pool_alloc(ctx->device, total, &pop->d_blob, &pop->blob_cap);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(ctx, DEX_ERR_NOMEM, std::string("population storage: ") + cudaGetErrorString(e)); }
    // one pinned staging image, one copy
    int rc = ensure_pinned(ctx, uploaded);
    if (rc) return rc;
    char* stage = static_cast<char*>(ctx->pinned);
    for (const Item& it : items) {
        *it.dst = static_cast<char*>(pop->d_blob) + it.off;
        if (it.bytes) std::memcpy(stage + it.off, it.src, it.bytes);
    }
    if (uploaded) CU(ctx, cudaMemcpyAsync(pop->d_blob, stage, uploaded, cudaMemcpyHostToDevice, ctx->stream));
    return DEX_OK;
}

// tree-index ranges with balanced tape length
int chunk_table(dex_ctx* ctx, dex_population* pop, int32_t n_chunks, const int32_t** out) {
    for (int k = 0; k < 2; ++k)
        if (pop->chunk_key[k] == n_chunks) { *out = pop->d_chunk_slot[k]; return DEX_OK; }
    std::vector<int32_t>& tab = pop->chunk_tables_host[n_chunks];
    if (tab.empty()) {
        const PackedPopulation& h = *pop->h.folded;
        tab.assign((size_t)n_chunks + 1, 0);
        // cost of trees [0, t) = tape instructions + a fixed per-tree cost of one
        const std::vector<int64_t>& toff = h.tape_off;
        auto cum = [&](int64_t t) { return toff[(size_t)t] + t; };
        const int64_t total = cum(h.n_trees);
        int64_t t = 0;
        for (int32_t c = 1; c < n_chunks; ++c) {
            const int64_t target = (total * c) / n_chunks;
            while (t < h.n_trees && cum(t) < target) ++t;
            tab[c] = (int32_t)t;
        }
        tab[n_chunks] = (int32_t)h.n_trees;
    }
    const int slot = pop->chunk_next;
    pop->chunk_next ^= 1;
    // pageable source: the runtime stages it before returning, and the copy is ordered on the
    // stream behind any kernel that still reads the slot's previous table
    CU(ctx, cudaMemcpyAsync(pop->d_chunk_slot[slot], tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                            ctx->stream));
    pop->chunk_key[slot] = n_chunks;
    *out = pop->d_chunk_slot[slot];
    return DEX_OK;
}

int check_eval_args(dex_ctx* ctx, const dex_population* pop, const void* X, int32_t F, int64_t N,
                    int64_t ldx, const void* out, int64_t ldo, const uint8_t* ok) {
    if (!pop) return set_err(ctx, DEX_ERR_INVALID, "null population");
    if (pop->device != ctx->device) return set_err(ctx, DEX_ERR_INVALID, "population was packed on another device");
    if (F < 0 || N < 0) return set_err(ctx, DEX_ERR_INVALID, "negative size");
    if (N > 0 && (!X && F > 0)) return set_err(ctx, DEX_ERR_INVALID, "null X");
    if (ldx < F) return set_err(ctx, DEX_ERR_INVALID, "ldx < nfeatures");
    if (out && ldo < N) return set_err(ctx, DEX_ERR_INVALID, "ldo < nsamples");
    if (!ok && pop->h.n_trees > 0) return set_err(ctx, DEX_ERR_INVALID, "null ok");
    if (pop->h.max_feature >= F)
        return set_err(ctx, DEX_ERR_RANGE,
                       "population uses feature " + std::to_string(pop->h.max_feature + 1) +
                           " (1-based) but X has only " + std::to_string(F) + " rows");
    return DEX_OK;
}

// Constant folding is a function of the constants only: run the scalar tape once per change of
// the constants (launch_fold) instead of in every call's prepass.  Returns the per-tree outcome.
int ensure_folded(dex_ctx* ctx, dex_population* pop, int rule, const uint8_t** fold_ok) {
    *fold_ok = nullptr;
    const PackedPopulation& f = *pop->h.folded;
    if (f.seg.empty()) return DEX_OK;          // nothing to fold: the prepass presets ok[] = 1
    if (!pop->fold_valid[rule]) {
        cudaError_t e = launch_fold(f.dtype, rule == 1, pop->d_ftape, pop->d_ctape, pop->d_seg, pop->d_seg_off,
                                    f.n_trees, pop->d_fold_ok[rule], ctx->stream);
        if (e != cudaSuccess) return cuda_err(ctx, e, "constant folding");
        ctx->launches += 1;
        // the result is reused by later calls, possibly issued on another stream
        // (dex_ctx_set_stream): complete it now — this happens once per change of the constants
        CU(ctx, cudaStreamSynchronize(ctx->stream));
        pop->fold_valid[rule] = true;
    }
    *fold_ok = pop->d_fold_ok[rule];
    return DEX_OK;
}

// Zero samples: nothing to launch, but the `complete` flags are still an output.  The reference
// validates an empty array as is_valid(sum(())) = true, so a tree is complete unless one of its
// folded constant subtrees is not (/root/reference/src/Evaluate.jl:347-354).
// rule: 0 evaluation, 1 gradient (folded launches only), -1 no folding on this path
int preset_ok_empty(dex_ctx* ctx, dex_population* pop, int rule, uint8_t* ok) {
    const uint8_t* fold_ok = nullptr;
    if (rule >= 0) {
        int rc = ensure_folded(ctx, pop, rule, &fold_ok);
        if (rc) return rc;
    }
    const size_t P = (size_t)pop->h.n_trees;
    if (fold_ok) CU(ctx, cudaMemcpyAsync(ok, fold_ok, P, cudaMemcpyDeviceToDevice, ctx->stream));
    else CU(ctx, cudaMemsetAsync(ok, 1, P, ctx->stream));
    return DEX_OK;
}

int run_eval(dex_ctx* ctx, const dex_population* cpop, const void* X, int32_t F, int64_t N,
             int64_t ldx, void* out, int64_t ldo, uint8_t* ok, int eval_flags, const void* params,
             int32_t n_params, int32_t n_classes, const int32_t* classes, const void* y,
             const void* w, double* loss_partial, int64_t* n_tiles_out, int n_slices = 1,
             void* out_host = nullptr, int64_t ldo_host = 0, uint8_t* ok_host = nullptr) {
    dex_population* pop = const_cast<dex_population*>(cpop);
    const PackedPopulation& h = *pop->h.folded;   // evaluation runs the folded image
    if (h.n_trees == 0) return DEX_OK;
    if (N == 0) return preset_ok_empty(ctx, pop, 0, ok);
    int threads;
    size_t smem;
    const int32_t front_rows = h.max_stack + h.n_param_rows;
    const int wide = eval_wide_mode((eval_flags & DEX_EVAL_EARLY_EXIT) != 0, params != nullptr, y != nullptr);   // as launch_eval decides
    const int64_t n_tiles = eval_num_tiles(h.dtype, F, front_rows, N, &threads, &smem, wide);
    if (n_tiles_out) *n_tiles_out = n_tiles;
    if (smem > 227 * 1024)
        return set_err(ctx, DEX_ERR_UNSUPPORTED,
                       "nfeatures + stack rows = " + std::to_string(F + front_rows) +
                           " do not fit in shared memory");
    // Chunking policy.  (a) enough CTAs for ~4 waves of 8 resident CTAs per SM; (b) chunks of at
    // most ~chunk_instr tape instructions: CTAs are dispatched tile-fastest, so with short
    // chunks the CTAs running at any moment write the rows of only a few dozen trees — the
    // output pages they touch stay within TLB reach (with 10^3 trees per chunk every store
    // after a tree switch is a page walk, and a 42 GB result costs +70 % time).
    const int64_t want = (int64_t)ctx->sm_count * 8 * 4;
    int64_t n_chunks = std::max<int64_t>(1, (want + n_tiles - 1) / n_tiles);
    int64_t chunk_instr = y ? 256 : 128;
    // long tapes (C4's depth-12 trees average 43 instructions, 71 % of them abandoned early): shorter chunks
    // balance better across CTAs — C4 shard 13.07 ms at 128, 12.83 at 96, 12.76 at 64; populations of short
    // tapes want the longer ones (C6 37.16 ms at 128, 37.73 at 64)
    if (!y && (int64_t)h.tape.size() >= 32 * h.n_trees) chunk_instr = 64;
    if (const char* env = getenv("DEXB200_CHUNK_INSTR")) chunk_instr = std::max<int64_t>(1, atoll(env));
    n_chunks = std::max<int64_t>(n_chunks, ((int64_t)h.tape.size() + chunk_instr - 1) / chunk_instr);
    n_chunks = std::min<int64_t>(n_chunks, std::min<int64_t>(h.n_trees, 65535));
    EvalArgs a{};
    a.dtype = h.dtype;
    a.tape = pop->d_ftape;
    a.tape_off = pop->d_ftape_off;
    int rc = ensure_folded(ctx, pop, 0, &a.fold_ok);
    if (rc) return rc;
    a.n_trees = h.n_trees;
    if ((rc = chunk_table(ctx, pop, (int32_t)n_chunks, &a.chunk_start))) return rc;
    a.n_chunks = (int32_t)n_chunks;
    a.max_stack = h.max_stack;
    a.n_param_rows = h.n_param_rows;
    if ((rc = ensure_xt(ctx, eval_xt_bytes(h.dtype, F, front_rows, N, wide)))) return rc;
    a.X = X; a.F = F; a.N = N; a.ldx = ldx; a.xt = ctx->xt;
    a.out = out; a.ldo = ldo; a.ok = ok;
    a.early_exit = (eval_flags & DEX_EVAL_EARLY_EXIT) ? 1 : 0;
    a.params = params; a.n_params = n_params; a.n_classes = n_classes; a.classes = classes;
    a.y = y; a.w = w; a.loss_partial = loss_partial;
    // CTA barrier every sync_tree-th tree (a power of two, 0 = never): off.  With the single-chain loss
    // epilogue of earlier versions the warps of a CTA drifted apart without one (+7 %); with four
    // accumulation chains the barrier only costs: C6-loss every tree 40.96 ms, every 2nd 40.30, 4th 39.75,
    // 8th 39.65, 32nd 39.22, never 39.13.  Store path (C2 / C4 shard / C6): never 0.2888 / 14.79 / 37.81,
    // every 8th 0.2887 / 14.86 / 38.25, every 4th 0.353 / 14.96 / 38.64.
    {
        auto pow2 = [](const char* env, int dflt) {
            if (!env) return dflt;
            int v = atoi(env), p = 0;
            while (v > 1) { v >>= 1; ++p; }
            return atoi(env) <= 0 ? 0 : 1 << p;
        };
        a.sync_tree = y ? pow2(getenv("DEXB200_LOSS_SYNC"), 0) : pow2(getenv("DEXB200_SYNC_TREE"), 0);
    }
    int launches = 0;
    if (!out_host) {
        cudaError_t e = launch_eval(a, ctx->stream, ctx->sm_count, &launches);
        ctx->launches += launches;
        if (e != cudaSuccess) return cuda_err(ctx, e, "eval kernel launch");
        return DEX_OK;
    }
    // Host-result pipeline: the population is evaluated in slices of consecutive tree chunks;
    // the device->host copy of slice s (copy stream) overlaps the kernel of slice s+1.
    // DEX_EVAL_SKIP_INCOMPLETE: the flags of a slice come back first and only the rows of complete
    // trees are transferred (runs of consecutive complete trees, one strided copy each) — the
    // others are unspecified under early exit, and device->host bandwidth is what the host entry
    // point is bound by.
    const bool skip = ok_host && (eval_flags & DEX_EVAL_SKIP_INCOMPLETE) && a.early_exit;
    const std::vector<int32_t>& tab = pop->chunk_tables_host[(int32_t)n_chunks];
    const size_t es = h.dtype == DEX_F32 ? 4 : 8;
    n_slices = (int)std::max<int64_t>(1, std::min<int64_t>(n_slices, n_chunks));
    auto slice_trees = [&](int sl, int64_t& t_lo, int64_t& t_hi) {
        t_lo = tab[(size_t)(n_chunks * sl / n_slices)];
        t_hi = tab[(size_t)(n_chunks * (sl + 1) / n_slices)];
    };
    auto launch_slice = [&](int sl) -> int {
        const int64_t c0 = n_chunks * sl / n_slices, c1 = n_chunks * (sl + 1) / n_slices;
        EvalArgs b = a;
        b.chunk_start = a.chunk_start + c0;
        b.n_chunks = (int32_t)(c1 - c0);
        b.skip_prepass = sl > 0;
        cudaError_t e = launch_eval(b, ctx->stream, ctx->sm_count, &launches);
        if (e != cudaSuccess) return cuda_err(ctx, e, "eval kernel launch");
        int64_t t_lo, t_hi;
        slice_trees(sl, t_lo, t_hi);
        if (skip && t_hi > t_lo)
            CU(ctx, cudaMemcpyAsync(ok_host + t_lo, ok + t_lo, (size_t)(t_hi - t_lo), cudaMemcpyDeviceToHost, ctx->stream));
        CU(ctx, cudaEventRecord(ctx->ev[sl & 1], ctx->stream));
        return DEX_OK;
    };
    auto copy_rows = [&](int64_t t0, int64_t t1) -> int {
        CU(ctx, cudaMemcpy2DAsync(static_cast<char*>(out_host) + (size_t)t0 * (size_t)ldo_host * es,
                                  (size_t)ldo_host * es,
                                  static_cast<const char*>(out) + (size_t)t0 * (size_t)ldo * es,
                                  (size_t)ldo * es, (size_t)N * es, (size_t)(t1 - t0),
                                  cudaMemcpyDeviceToHost, ctx->copy_stream));
        return DEX_OK;
    };
    rc = launch_slice(0);
    for (int sl = 0; sl < n_slices && rc == DEX_OK; ++sl) {
        if (sl + 1 < n_slices && (rc = launch_slice(sl + 1))) break;     // keeps the device busy while the host waits
        int64_t t_lo, t_hi;
        slice_trees(sl, t_lo, t_hi);
        if (!skip) {
            CU(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[sl & 1], 0));
            if (t_hi > t_lo) rc = copy_rows(t_lo, t_hi);
            continue;
        }
        CU(ctx, cudaEventSynchronize(ctx->ev[sl & 1]));     // the flags of this slice are on the host
        for (int64_t t = t_lo; t < t_hi && rc == DEX_OK;) {
            if (!ok_host[t]) { ++t; continue; }
            int64_t e = t + 1;
            while (e < t_hi && ok_host[e]) ++e;
            rc = copy_rows(t, e);
            t = e;
        }
    }
    ctx->launches += launches;
    return rc;
}

}  // namespace

// ---- library -------------------------------------------------------------------------
extern "C" {

int dex_abi_version(void) { return DEXB200_ABI_VERSION; }

const char* dex_strerror(int code) {
    switch (code) {
        case DEX_OK: return "ok";
        case DEX_ERR_INVALID: return "invalid argument or malformed tree";
        case DEX_ERR_NOMEM: return "out of memory";
        case DEX_ERR_CUDA: return "CUDA failure or no device";
        case DEX_ERR_UNSUPPORTED: return "unsupported (operator without device implementation, tree too deep, ...)";
        case DEX_ERR_RANGE: return "feature / parameter / class index out of range";
        default: return "unknown error";
    }
}

int dex_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int dex_opcode_from_name(const char* name, int degree) {
    if (!name) return -1;
    for (const OpInfo& o : kOps)
        if (o.degree == degree && std::strcmp(o.name, name) == 0) return o.code;
    for (const OpInfo& o : kOps)
        if (o.degree == degree && alias_match(o.aliases, name)) return o.code;
    return -1;
}
const char* dex_opcode_name(int opcode) {
    const OpInfo* o = find_op(opcode);
    return o ? o->name : nullptr;
}
int dex_opcode_degree(int opcode) {
    const OpInfo* o = find_op(opcode);
    return o ? o->degree : 0;
}

// ---- context ---------------------------------------------------------------------------
int dex_ctx_create(int device, dex_ctx** out) {
    if (!out) return DEX_ERR_INVALID;
    *out = nullptr;
    dex_ctx* ctx = new (std::nothrow) dex_ctx();
    if (!ctx) return DEX_ERR_NOMEM;
    ctx->device = device;
    if (device >= 0) {
        int n = dex_device_count();
        if (device >= n) { delete ctx; return DEX_ERR_CUDA; }
        if (cudaSetDevice(device) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev[1], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            delete ctx;
            return DEX_ERR_CUDA;
        }
        ctx->stream = ctx->own_stream;
        cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    }
    *out = ctx;
    return DEX_OK;
}

int dex_ctx_destroy(dex_ctx* ctx) {
    if (!ctx) return DEX_OK;
    if (ctx->device >= 0) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->scratch) cudaFree(ctx->scratch);
        if (ctx->xt) cudaFree(ctx->xt);
        if (ctx->dev_io) cudaFree(ctx->dev_io);
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
        if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
        if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
        if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    }
    delete ctx;
    return DEX_OK;
}

int dex_ctx_set_stream(dex_ctx* ctx, void* stream) {
    if (!ctx) return DEX_ERR_INVALID;
    cudaStream_t next = static_cast<cudaStream_t>(stream);
    int rc = order_after_previous(ctx, next);
    if (rc) return rc;
    ctx->stream = next;
    return DEX_OK;
}

int dex_ctx_use_own_stream(dex_ctx* ctx) {
    if (!ctx) return DEX_ERR_INVALID;
    int rc = order_after_previous(ctx, ctx->own_stream);
    if (rc) return rc;
    ctx->stream = ctx->own_stream;
    return DEX_OK;
}

int dex_ctx_synchronize(dex_ctx* ctx) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return DEX_OK;
}

const char* dex_last_error(const dex_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int64_t dex_ctx_launch_count(const dex_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dex_eval_launch_info(const dex_population* pop, int32_t nfeatures, int64_t nsamples, int eval_flags,
                         int loss, int parametric, int32_t* threads, int64_t* smem_bytes, int64_t* n_tiles,
                         int32_t* smem_rows) {
    if (!pop || nfeatures < 0 || nsamples < 0) return DEX_ERR_INVALID;
    const PackedPopulation& h = *pop->h.folded;
    int th = 0, rows = 0;
    size_t smem = 0;
    const int wide = eval_wide_mode((eval_flags & DEX_EVAL_EARLY_EXIT) != 0, parametric != 0, loss != 0);
    const int64_t tiles = eval_num_tiles(h.dtype, nfeatures, h.max_stack + h.n_param_rows,
                                         std::max<int64_t>(nsamples, 1), &th, &smem, wide, &rows);
    if (threads) *threads = th;
    if (smem_bytes) *smem_bytes = (int64_t)smem;
    if (n_tiles) *n_tiles = tiles;
    if (smem_rows) *smem_rows = rows;
    return DEX_OK;
}

// ---- operators -------------------------------------------------------------------------
int dex_optable_create(const int32_t* opcodes, const int32_t* degree_offsets, int max_degree,
                       dex_optable** out) {
    if (!out || !degree_offsets || max_degree < 0 || max_degree > DEX_MAX_DEGREE) return DEX_ERR_INVALID;
    dex_optable* t = new (std::nothrow) dex_optable();
    if (!t) return DEX_ERR_NOMEM;
    for (int d = 0; d < max_degree; ++d) {
        for (int32_t i = degree_offsets[d]; i < degree_offsets[d + 1]; ++i) {
            const OpInfo* o = opcodes ? find_op(opcodes[i]) : nullptr;
            if (!o || o->degree != d + 1) { delete t; return DEX_ERR_UNSUPPORTED; }
            t->t.ops[d].push_back(opcodes[i]);
        }
        if (t->t.ops[d].size() > 256) { delete t; return DEX_ERR_INVALID; }
    }
    *out = t;
    return DEX_OK;
}
int dex_optable_destroy(dex_optable* t) {
    delete t;
    return DEX_OK;
}

// ---- populations -----------------------------------------------------------------------
int dex_population_pack(dex_ctx* ctx, const dex_optable* ops, const dex_node* nodes,
                        const int64_t* offsets, int64_t n_trees, int dtype, int pack_flags,
                        dex_population** out) {
    DEX_RANGE("dex_population_pack");
    if (!ctx || !out) return DEX_ERR_INVALID;
    *out = nullptr;
    if (!ops || !offsets || n_trees < 0 || (n_trees > 0 && !nodes)) return set_err(ctx, DEX_ERR_INVALID, "null argument");
    if (dtype != DEX_F32 && dtype != DEX_F64) return set_err(ctx, DEX_ERR_INVALID, "dtype must be DEX_F32 or DEX_F64");
    if (offsets[0] < 0) return set_err(ctx, DEX_ERR_INVALID, "offsets[0] is negative");
    for (int64_t t = 0; t < n_trees; ++t)
        if (offsets[t + 1] <= offsets[t])
            return set_err(ctx, DEX_ERR_INVALID, "tree " + std::to_string(t) + ": offsets must be strictly increasing (empty tree or overlapping records)");
    dex_population* pop = new (std::nothrow) dex_population();
    if (!pop) return set_err(ctx, DEX_ERR_NOMEM, "out of host memory");
    std::string err;
    int rc = flatten_population(ops->t, nodes, offsets, n_trees, dtype, pack_flags, pop->h, err);
    if (rc) { delete pop; return set_err(ctx, rc, err); }
    pop->device = ctx->device;
    if (ctx->device >= 0) {
        rc = ensure_device(ctx);
        if (!rc) rc = upload_population(ctx, pop);
        if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = set_err(ctx, DEX_ERR_CUDA, "tape upload failed");
        if (rc) { dex_population_destroy(pop); return rc; }
    }
    *out = pop;
    return DEX_OK;
}

int dex_population_destroy(dex_population* pop) {
    if (!pop) return DEX_OK;
    if (pop->device >= 0) {
        cudaSetDevice(pop->device);
        // as cudaFree would: nothing on the device still reads the tapes once this returns, so the
        // block can be handed to the next population straight away
        if (pop->d_blob) cudaDeviceSynchronize();
        pool_free(pop->device, pop->d_blob, pop->blob_cap);
    }
    delete pop;
    return DEX_OK;
}

int dex_population_get_info(const dex_population* pop, dex_population_info* info) {
    if (!pop || !info) return DEX_ERR_INVALID;
    info->n_trees = pop->h.n_trees;
    info->n_nodes = pop->h.n_nodes;
    info->n_instructions = (int64_t)pop->h.tape.size();
    info->n_constants = pop->h.n_constants;
    info->max_stack = pop->h.max_stack;
    info->max_feature = pop->h.max_feature;
    info->max_parameter = pop->h.max_parameter;
    info->dtype = pop->h.dtype;
    info->n_generic = pop->h.n_generic;
    info->n_checks = pop->h.n_checks;
    info->n_folded_instructions = (int64_t)pop->h.folded->tape.size();
    info->n_scalar_instructions = (int64_t)pop->h.folded->ctape.size();
    info->n_folded_subtrees = (int64_t)pop->h.folded->seg.size() / 3;
    info->folded_max_stack = pop->h.folded->max_stack;
    return DEX_OK;
}

int dex_population_constant_counts(const dex_population* pop, int32_t* counts) {
    if (!pop || !counts) return DEX_ERR_INVALID;
    std::copy(pop->h.n_const_tree.begin(), pop->h.n_const_tree.end(), counts);
    return DEX_OK;
}

const char* dex_handler_name(int h) {
    static const char* names[] = {
        "GENERIC", "LOAD_R", "LOAD_C",
#define X(S) #S "_A", #S "_R",
        DEX_FAST_UNARY(X)
#undef X
#define X(S) #S "_AR", #S "_AC", #S "_RR", #S "_RC",
        DEX_FAST_BIN_COMM(X)
#undef X
#define X(S) #S "_AR", #S "_RA", #S "_AC", #S "_CA", #S "_RR", #S "_RC", #S "_CR",
        DEX_FAST_BIN_NC(X)
#undef X
        "KEEP",
    };
    static_assert(sizeof(names) / sizeof(names[0]) == H__COUNT, "handler name table out of sync");
    return (h >= 0 && h < (int)H__COUNT) ? names[h] : nullptr;
}

// debugging / tests: copy out the host image of the evaluation tape
int64_t dex_population_copy_tape(const dex_population* pop, void* instrs, int64_t capacity,
                                 int64_t* offsets /* n_trees+1 */) {
    if (!pop) return DEX_ERR_INVALID;
    const int64_t n = (int64_t)pop->h.tape.size();
    if (instrs && capacity >= n) std::memcpy(instrs, pop->h.tape.data(), (size_t)n * sizeof(Instr));
    if (offsets) std::copy(pop->h.tape_off.begin(), pop->h.tape_off.end(), offsets);
    return n;
}

// the folded image: main tape, scalar tape, segment table (3 int64 per folded subtree)
int dex_population_copy_folded(const dex_population* pop, void* instrs, int64_t* offsets,
                               void* scalar_instrs, int64_t* segs, int64_t* seg_offsets) {
    if (!pop) return DEX_ERR_INVALID;
    const PackedPopulation& f = *pop->h.folded;
    if (instrs) std::memcpy(instrs, f.tape.data(), f.tape.size() * sizeof(Instr));
    if (offsets) std::copy(f.tape_off.begin(), f.tape_off.end(), offsets);
    if (scalar_instrs) std::memcpy(scalar_instrs, f.ctape.data(), f.ctape.size() * sizeof(Instr));
    if (segs) std::copy(f.seg.begin(), f.seg.end(), segs);
    if (seg_offsets) std::copy(f.seg_off.begin(), f.seg_off.end(), seg_offsets);
    return DEX_OK;
}

int dex_population_get_constants(dex_ctx* ctx, const dex_population* pop, void* values_host,
                                 int64_t n_values) {
    if (!ctx || !pop || (!values_host && n_values > 0)) return DEX_ERR_INVALID;
    if (n_values != pop->h.n_constants) return set_err(ctx, DEX_ERR_INVALID, "constant count mismatch");
    for (int64_t k = 0; k < n_values; ++k) {
        const Instr& ins = pop->h.tape[(size_t)pop->h.const_pos[(size_t)k]];
        if (pop->h.dtype == DEX_F32) std::memcpy(static_cast<float*>(values_host) + k, &ins.c_lo, 4);
        else { uint64_t u = ((uint64_t)ins.c_hi << 32) | ins.c_lo; std::memcpy(static_cast<double*>(values_host) + k, &u, 8); }
    }
    return DEX_OK;
}

int dex_population_set_constants(dex_ctx* ctx, dex_population* pop, const void* values_host,
                                 int64_t n_values) {
    DEX_RANGE("dex_population_set_constants");
    if (!ctx || !pop || (!values_host && n_values > 0)) return DEX_ERR_INVALID;
    if (n_values != pop->h.n_constants) return set_err(ctx, DEX_ERR_INVALID, "constant count mismatch");
    const size_t es = pop->h.dtype == DEX_F32 ? 4 : 8;
    // host image
    for (int64_t k = 0; k < n_values; ++k) {
        Instr& ins = pop->h.tape[(size_t)pop->h.const_pos[(size_t)k]];
        if (es == 4) { std::memcpy(&ins.c_lo, static_cast<const float*>(values_host) + k, 4); ins.c_hi = 0; }
        else { uint64_t u; std::memcpy(&u, static_cast<const double*>(values_host) + k, 8); ins.c_lo = (uint32_t)u; ins.c_hi = (uint32_t)(u >> 32); }
        PackedPopulation& f = *pop->h.folded;
        const int64_t fp = f.const_pos[(size_t)k];
        Instr& fi = fp >= 0 ? f.tape[(size_t)fp] : f.ctape[(size_t)(-(fp + 1))];
        fi.c_lo = ins.c_lo;
        fi.c_hi = ins.c_hi;
    }
    if (pop->device < 0 || n_values == 0) return DEX_OK;
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = ensure_pinned(ctx, (size_t)n_values * es))) return rc;
    if ((rc = ensure_dev_io(ctx, (size_t)n_values * es))) return rc;
    std::memcpy(ctx->pinned, values_host, (size_t)n_values * es);
    CU(ctx, cudaMemcpyAsync(ctx->dev_io, ctx->pinned, (size_t)n_values * es, cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = launch_scatter_constants(pop->h.dtype, pop->d_tape, nullptr, pop->d_const_pos, ctx->dev_io,
                                             n_values, ctx->stream);
    if (e == cudaSuccess)
        e = launch_scatter_constants(pop->h.dtype, pop->d_ftape, pop->d_ctape, pop->d_fconst_pos, ctx->dev_io,
                                     n_values, ctx->stream);
    if (e != cudaSuccess) return cuda_err(ctx, e, "scatter constants");
    ctx->launches += 2;
    pop->fold_valid[0] = pop->fold_valid[1] = false;   // the folded constants are stale now
    CU(ctx, cudaStreamSynchronize(ctx->stream));  // pinned staging is reused by later calls
    return DEX_OK;
}

// ---- evaluation ------------------------------------------------------------------------
int dex_eval(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
             int64_t nsamples, int64_t ldx, void* out_dev, int64_t ldo, uint8_t* ok_dev,
             int eval_flags) {
    DEX_RANGE("dex_eval");
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = check_eval_args(ctx, pop, X_dev, nfeatures, nsamples, ldx, out_dev, ldo, ok_dev))) return rc;
    if (!out_dev && nsamples > 0 && pop->h.n_trees > 0) return set_err(ctx, DEX_ERR_INVALID, "null out");
    if (pop->h.max_parameter >= 0) return set_err(ctx, DEX_ERR_INVALID, "population has parameter leaves: use dex_eval_parametric");
    return finish_call(ctx, run_eval(ctx, pop, X_dev, nfeatures, nsamples, ldx, out_dev, ldo, ok_dev, eval_flags,
                                     nullptr, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr));
}

int dex_eval_parametric(dex_ctx* ctx, const dex_population* pop, const void* X_dev,
                        int32_t nfeatures, int64_t nsamples, int64_t ldx, const void* params_dev,
                        int32_t n_params, int32_t n_classes, const int32_t* classes_dev,
                        void* out_dev, int64_t ldo, uint8_t* ok_dev, int eval_flags) {
    DEX_RANGE("dex_eval_parametric");
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = check_eval_args(ctx, pop, X_dev, nfeatures, nsamples, ldx, out_dev, ldo, ok_dev))) return rc;
    if (!out_dev && nsamples > 0 && pop->h.n_trees > 0) return set_err(ctx, DEX_ERR_INVALID, "null out");
    if (n_params < 0 || n_classes < 1 || !classes_dev || (n_params > 0 && !params_dev))
        return set_err(ctx, DEX_ERR_INVALID, "bad parameter arguments");
    if (pop->h.n_param_rows > n_params)
        return set_err(ctx, DEX_ERR_RANGE, "population uses (or was packed for) " + std::to_string(pop->h.n_param_rows) +
                                               " parameters but only " + std::to_string(n_params) + " were passed");
    // a dummy non-null params pointer keeps the parametric kernel selected when n_params == 0
    return finish_call(ctx, run_eval(ctx, pop, X_dev, nfeatures, nsamples, ldx, out_dev, ldo, ok_dev, eval_flags,
                                     params_dev ? params_dev : X_dev, n_params, n_classes, classes_dev, nullptr,
                                     nullptr, nullptr, nullptr));
}

int dex_eval_loss(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, const void* y_dev, const void* weights_dev,
                  double* loss_dev, uint8_t* ok_dev, int eval_flags) {
    DEX_RANGE("dex_eval_loss");
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = check_eval_args(ctx, pop, X_dev, nfeatures, nsamples, ldx, nullptr, 0, ok_dev))) return rc;
    if (!y_dev || !loss_dev) return set_err(ctx, DEX_ERR_INVALID, "null y / loss");
    if (pop->h.max_parameter >= 0) return set_err(ctx, DEX_ERR_INVALID, "population has parameter leaves");
    if (pop->h.n_trees == 0) return DEX_OK;
    int threads; size_t smem;
    const int64_t n_tiles = eval_num_tiles(pop->h.dtype, nfeatures, pop->h.folded->max_stack + pop->h.n_param_rows, std::max<int64_t>(nsamples, 1), &threads, &smem,
                                           eval_wide_mode((eval_flags & DEX_EVAL_EARLY_EXIT) != 0, false, true));
    // scratch: [sum of weights (256 B slot)] [partial sums: one per (tile, warp slot, tree); warps a
    // smaller CTA does not have leave their zero]
    constexpr int64_t kWarpSlots = 8;   // DEX_MAX_THREADS / 32 (dex_eval.cu)
    const size_t partial_bytes = (size_t)n_tiles * kWarpSlots * (size_t)pop->h.n_trees * sizeof(double);
    if ((rc = ensure_scratch(ctx, 256 + partial_bytes))) return rc;
    double* wsum = static_cast<double*>(ctx->scratch);
    double* partial = reinterpret_cast<double*>(static_cast<char*>(ctx->scratch) + 256);
    if (weights_dev && nsamples > 0) {
        cudaError_t e = launch_weight_sum(pop->h.dtype, weights_dev, nsamples, wsum, ctx->stream);
        if (e != cudaSuccess) return cuda_err(ctx, e, "weight sum");
        ctx->launches += 1;
    }
    if (threads < 256 && nsamples > 0) CU(ctx, cudaMemsetAsync(partial, 0, partial_bytes, ctx->stream));
    if ((rc = run_eval(ctx, pop, X_dev, nfeatures, nsamples, ldx, nullptr, 0, ok_dev, eval_flags, nullptr, 0,
                       0, nullptr, y_dev, weights_dev, partial, nullptr)))
        return rc;
    cudaError_t e = launch_loss_grad_reduce(partial, nsamples > 0 ? n_tiles * kWarpSlots : 0, pop->h.n_trees, pop->h.n_trees,
                                            nsamples > 0 ? 1.0 / (double)nsamples : 0.0,
                                            (weights_dev && nsamples > 0) ? wsum : nullptr, loss_dev, nullptr,
                                            ctx->stream);
    if (e != cudaSuccess) return cuda_err(ctx, e, "loss reduce");
    ctx->launches += 1;
    return finish_call(ctx, DEX_OK);
}

int dex_grad_offsets(const dex_population* pop, int32_t nfeatures, int64_t nsamples, int mode,
                     int64_t* offsets_host) {
    if (!pop || !offsets_host || mode < 0 || mode > 2 || nfeatures < 0 || nsamples < 0) return DEX_ERR_INVALID;
    int64_t acc = 0;
    for (int64_t t = 0; t < pop->h.n_trees; ++t) {
        offsets_host[t] = acc;
        const int64_t nc = pop->h.n_const_tree[(size_t)t];
        const int64_t G = mode == DEX_GRAD_FEATURES ? nfeatures : mode == DEX_GRAD_CONSTANTS ? nc : nfeatures + nc;
        acc += G * nsamples;
    }
    offsets_host[pop->h.n_trees] = acc;
    return DEX_OK;
}

struct LossSpec {            // fused loss + gradient of the loss (dex_eval_loss_grad)
    const void* y;
    const void* w;
    double* loss;
    double* grad;
};

struct ParamSpec {           // ParametricExpression arguments of a gradient call
    const void* params;
    int32_t n_params, n_classes;
    const int32_t* classes;
};

// FX: rows of the caller's X.  With parameters the leaf rows of the launch are the n_param_rows
// gathered parameter rows followed by the FX features (F below counts both).
static int run_grad(dex_ctx* ctx, const dex_population* cpop, const void* X, int32_t FX, int64_t N,
                    int64_t ldx, int mode, int32_t direction, void* out, int64_t ldo, void* grad,
                    const int64_t* grad_offsets_host, uint8_t* ok, const LossSpec* ls = nullptr,
                    const ParamSpec* ps = nullptr) {
    dex_population* pop = const_cast<dex_population*>(cpop);
    const PackedPopulation& h = pop->h;
    if (!ps && h.max_parameter >= 0)
        return set_err(ctx, DEX_ERR_INVALID, "population has parameter leaves: use dex_eval_grad_parametric");
    if (ps && ps->n_params != h.n_param_rows)
        return set_err(ctx, DEX_ERR_INVALID,
                       "gradients of a parametric population need one row per parameter: pack it with "
                       "DEX_PACK_PARAM_ROWS(" + std::to_string(ps->n_params) + ") (it has " +
                           std::to_string(h.n_param_rows) + " parameter rows)");
    const int32_t F = FX + (ps ? h.n_param_rows : 0);
    if (h.n_trees == 0) return DEX_OK;
    if (N == 0) return preset_ok_empty(ctx, pop, mode == DEX_GRAD_FEATURES && F > 0 ? 1 : -1, ok);
    int Gmax = 1;
    if (mode >= 0) {
        int32_t ncmax = 0;
        for (int32_t c : h.n_const_tree) ncmax = std::max(ncmax, c);
        Gmax = mode == DEX_GRAD_FEATURES ? F : mode == DEX_GRAD_CONSTANTS ? ncmax : F + ncmax;
    }
    // d/dX only (variable = Val(true)): constant subtrees have zero gradient, so the launch runs
    // the constant-folded tape like dex_eval; the prepass evaluates the folded subtrees with the
    // gradient path's validity rule (every value and every partial finite, dex_fold.cuh).
    // Gradients w.r.t. constants and eval_diff run the unfolded tape.
    const bool folded = mode == DEX_GRAD_FEATURES && F > 0;
    const PackedPopulation& img = folded ? *h.folded : h;
    const int64_t n_tiles = grad_num_tiles(h.dtype, F, img.max_stack, Gmax, N, ls != nullptr);
    const int64_t want = (int64_t)ctx->sm_count * 8 * 4;
    int64_t n_chunks = std::max<int64_t>(1, (want + n_tiles - 1) / n_tiles);
    n_chunks = std::max<int64_t>(n_chunks, ((int64_t)img.tape.size() + 127) / 128);
    n_chunks = std::min<int64_t>(n_chunks, std::min<int64_t>(h.n_trees, 65535));
    const int32_t* chunks = nullptr;
    int rc = chunk_table(ctx, pop, (int32_t)n_chunks, &chunks);
    if (rc) return rc;
    if ((rc = ensure_xt(ctx, grad_xt_bytes(h.dtype, F, img.max_stack, Gmax, N, ls != nullptr)))) return rc;
    GradArgs a{};
    a.dtype = h.dtype;
    if (folded) {
        a.tape = pop->d_ftape; a.tape_off = pop->d_ftape_off;
        if ((rc = ensure_folded(ctx, pop, 1, &a.fold_ok))) return rc;
    } else {
        a.tape = pop->d_tape; a.tape_off = pop->d_tape_off;
    }
    a.const_ord = pop->d_const_ord;   // indexed like the unfolded tape; not read for d/dX
    a.const_off = pop->d_const_off;
    a.n_trees = h.n_trees; a.max_stack = img.max_stack;
    a.X = X; a.xt = ctx->xt; a.F = F; a.N = N; a.ldx = ldx; a.mode = mode; a.direction = direction;
    a.out = out; a.ldo = ldo; a.grad = grad; a.ok = ok; a.grad_off = nullptr;
    if (ps) {
        a.params = ps->params; a.classes = ps->classes;
        a.n_params = ps->n_params; a.n_classes = ps->n_classes; a.n_param_rows = h.n_param_rows;
    }
    double* wsum = nullptr;
    int64_t stride = 0;
    if (mode >= 0) {
        // scratch: [grad offsets (n_trees + 1) int64] [sum of weights] [per-tile partial sums]
        const size_t off_bytes = (((size_t)(h.n_trees + 1) * sizeof(int64_t)) + 255) & ~(size_t)255;
        size_t bytes = off_bytes;
        if (ls) {
            stride = h.n_trees + grad_offsets_host[h.n_trees];
            bytes += 256 + (size_t)n_tiles * (size_t)stride * sizeof(double);
        }
        if ((rc = ensure_scratch(ctx, bytes))) return rc;
        // pageable -> device: the runtime stages the source before returning, so the
        // caller's array may be reused immediately
        CU(ctx, cudaMemcpyAsync(ctx->scratch, grad_offsets_host, (size_t)(h.n_trees + 1) * sizeof(int64_t),
                                cudaMemcpyHostToDevice, ctx->stream));
        a.grad_off = static_cast<const int64_t*>(ctx->scratch);
        if (ls) {
            char* base = static_cast<char*>(ctx->scratch) + off_bytes;
            wsum = reinterpret_cast<double*>(base);
            a.partial = reinterpret_cast<double*>(base + 256);
            a.partial_stride = stride;
            a.y = ls->y; a.w = ls->w;
            // trees whose gradient has fewer entries than a pass leave nothing unwritten, but a
            // tree with G = 0 writes no gradient entry at all and entries of padded passes are
            // skipped: clear the buffer so that the reduction reads defined values
            CU(ctx, cudaMemsetAsync(a.partial, 0, (size_t)n_tiles * (size_t)stride * sizeof(double), ctx->stream));
            if (ls->w) {
                cudaError_t e2 = launch_weight_sum(h.dtype, ls->w, N, wsum, ctx->stream);
                if (e2 != cudaSuccess) return cuda_err(ctx, e2, "weight sum");
                ctx->launches += 1;
            }
        }
    }
    int launches = 0;
    cudaError_t e = launch_grad_ex(a, chunks, (int)n_chunks, Gmax, ctx->stream, &launches);
    ctx->launches += launches;
    if (e != cudaSuccess) return cuda_err(ctx, e, "grad kernel launch");
    if (ls) {
        e = launch_loss_grad_reduce(a.partial, n_tiles, stride, h.n_trees, 1.0 / (double)N, ls->w ? wsum : nullptr,
                                    ls->loss, ls->grad, ctx->stream);
        if (e != cudaSuccess) return cuda_err(ctx, e, "loss/gradient reduce");
        ctx->launches += 1;
    }
    return DEX_OK;
}

int dex_eval_loss_grad(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                       int64_t nsamples, int64_t ldx, const void* y_dev, const void* weights_dev, int mode,
                       double* loss_dev, double* grad_dev, const int64_t* grad_offsets_host, uint8_t* ok_dev) {
    DEX_RANGE("dex_eval_loss_grad");
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = check_eval_args(ctx, pop, X_dev, nfeatures, nsamples, ldx, nullptr, 0, ok_dev))) return rc;
    if (mode < 0 || mode > 2) return set_err(ctx, DEX_ERR_INVALID, "mode must be DEX_GRAD_CONSTANTS/FEATURES/BOTH");
    if (!y_dev || !loss_dev || !grad_offsets_host) return set_err(ctx, DEX_ERR_INVALID, "null y / loss / grad_offsets");
    if (!grad_dev && grad_offsets_host[pop->h.n_trees] > 0) return set_err(ctx, DEX_ERR_INVALID, "null grad");
    if (nsamples <= 0) return set_err(ctx, DEX_ERR_INVALID, "the loss needs at least one sample");
    LossSpec ls{y_dev, weights_dev, loss_dev, grad_dev};
    return finish_call(ctx, run_grad(ctx, pop, X_dev, nfeatures, nsamples, ldx, mode, 0, nullptr, 0, nullptr,
                                     grad_offsets_host, ok_dev, &ls));
}

int dex_eval_grad(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, int mode, void* out_dev, int64_t ldo,
                  void* grad_dev, const int64_t* grad_offsets_host, uint8_t* ok_dev) {
    DEX_RANGE("dex_eval_grad");
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = check_eval_args(ctx, pop, X_dev, nfeatures, nsamples, ldx, out_dev, ldo, ok_dev))) return rc;
    if (mode < 0 || mode > 2) return set_err(ctx, DEX_ERR_INVALID, "mode must be DEX_GRAD_CONSTANTS/FEATURES/BOTH");
    if ((!out_dev && nsamples > 0 && pop->h.n_trees > 0) || !grad_offsets_host) return set_err(ctx, DEX_ERR_INVALID, "null out / grad_offsets");
    if (!grad_dev && grad_offsets_host[pop->h.n_trees] > 0) return set_err(ctx, DEX_ERR_INVALID, "null grad");
    return finish_call(ctx, run_grad(ctx, pop, X_dev, nfeatures, nsamples, ldx, mode, 0, out_dev, ldo, grad_dev,
                                     grad_offsets_host, ok_dev));
}

int dex_eval_grad_parametric(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                             int64_t nsamples, int64_t ldx, const void* params_dev, int32_t n_params,
                             int32_t n_classes, const int32_t* classes_dev, int mode, void* out_dev,
                             int64_t ldo, void* grad_dev, const int64_t* grad_offsets_host, uint8_t* ok_dev) {
    DEX_RANGE("dex_eval_grad_parametric");
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = check_eval_args(ctx, pop, X_dev, nfeatures, nsamples, ldx, out_dev, ldo, ok_dev))) return rc;
    if (mode < 0 || mode > 2) return set_err(ctx, DEX_ERR_INVALID, "mode must be DEX_GRAD_CONSTANTS/FEATURES/BOTH");
    if ((!out_dev && nsamples > 0 && pop->h.n_trees > 0) || !grad_offsets_host) return set_err(ctx, DEX_ERR_INVALID, "null out / grad_offsets");
    if (!grad_dev && grad_offsets_host[pop->h.n_trees] > 0) return set_err(ctx, DEX_ERR_INVALID, "null grad");
    if (n_params < 0 || n_classes < 1 || !classes_dev || (n_params > 0 && !params_dev))
        return set_err(ctx, DEX_ERR_INVALID, "bad parameter arguments");
    ParamSpec ps{params_dev, n_params, n_classes, classes_dev};
    return finish_call(ctx, run_grad(ctx, pop, X_dev, nfeatures, nsamples, ldx, mode, 0, out_dev, ldo, grad_dev,
                                     grad_offsets_host, ok_dev, nullptr, &ps));
}

int dex_eval_diff(dex_ctx* ctx, const dex_population* pop, const void* X_dev, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, int32_t direction, void* out_dev, void* dout_dev,
                  int64_t ldo, uint8_t* ok_dev) {
    DEX_RANGE("dex_eval_diff");
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if ((rc = check_eval_args(ctx, pop, X_dev, nfeatures, nsamples, ldx, out_dev, ldo, ok_dev))) return rc;
    if ((!out_dev || !dout_dev) && nsamples > 0 && pop->h.n_trees > 0) return set_err(ctx, DEX_ERR_INVALID, "null out / dout");
    if (direction < 0) return set_err(ctx, DEX_ERR_RANGE, "direction must be a feature index");
    return finish_call(ctx, run_grad(ctx, pop, X_dev, nfeatures, nsamples, ldx, -1, direction, out_dev, ldo, dout_dev,
                                     nullptr, ok_dev));
}

// Enqueues everything of dex_eval_host (H2D of X, kernels, D2H of results and flags) without
// waiting for it; host_eval_wait() completes it.  Split so that one host thread can keep several
// devices busy (dex_shard_eval_host).
static int host_eval_enqueue(dex_ctx* ctx, const dex_population* pop, const void* X_host, int32_t nfeatures,
                             int64_t nsamples, int64_t ldx, void* out_host, int64_t ldo, uint8_t* ok_host,
                             int eval_flags) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (!pop || !out_host || !ok_host || (!X_host && nfeatures > 0 && nsamples > 0)) return set_err(ctx, DEX_ERR_INVALID, "null argument");
    if (pop->device != ctx->device) return set_err(ctx, DEX_ERR_INVALID, "population was packed on another device");
    if (nfeatures < 0 || nsamples < 0) return set_err(ctx, DEX_ERR_INVALID, "negative size");
    if (ldx < nfeatures) return set_err(ctx, DEX_ERR_INVALID, "ldx < nfeatures");
    if (ldo < nsamples) return set_err(ctx, DEX_ERR_INVALID, "ldo < nsamples");
    if (pop->h.max_parameter >= 0) return set_err(ctx, DEX_ERR_INVALID, "population has parameter leaves");
    if (pop->h.max_feature >= nfeatures)
        return set_err(ctx, DEX_ERR_RANGE,
                       "population uses feature " + std::to_string(pop->h.max_feature + 1) +
                           " (1-based) but X has only " + std::to_string(nfeatures) + " rows");
    const size_t es = pop->h.dtype == DEX_F32 ? 4 : 8;
    const int64_t P = pop->h.n_trees, N = nsamples;
    if (P == 0) return DEX_OK;
    if (N == 0) {   // no samples: only the flags are an output
        if ((rc = ensure_dev_io(ctx, (size_t)P))) return rc;
        uint8_t* dK0 = static_cast<uint8_t*>(ctx->dev_io);
        if ((rc = preset_ok_empty(ctx, const_cast<dex_population*>(pop), 0, dK0))) return rc;
        CU(ctx, cudaMemcpyAsync(ok_host, dK0, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
        return DEX_OK;
    }
    // device staging: X | out | ok
    const size_t xb = (size_t)ldx * (size_t)N * es;
    const size_t xoff = 0, ooff = (xb + 255) & ~(size_t)255;
    const size_t ob = (size_t)P * (size_t)N * es;
    const size_t koff = (ooff + ob + 255) & ~(size_t)255;
    if ((rc = ensure_dev_io(ctx, koff + (size_t)P))) return rc;
    char* base = static_cast<char*>(ctx->dev_io);
    void* dX = base + xoff;
    void* dO = base + ooff;
    uint8_t* dK = reinterpret_cast<uint8_t*>(base + koff);
    CU(ctx, cudaMemcpyAsync(dX, X_host, xb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = check_eval_args(ctx, pop, dX, nfeatures, N, ldx, dO, N, dK))) return rc;
    // results travel back slice by slice on the copy stream while the next slice computes
    const int n_slices = (size_t)P * (size_t)N * es >= ((size_t)8 << 20) ? 8 : 1;
    if ((rc = run_eval(ctx, pop, dX, nfeatures, N, ldx, dO, N, dK, eval_flags, nullptr, 0, 0, nullptr,
                       nullptr, nullptr, nullptr, nullptr, n_slices, out_host, ldo, ok_host)))
        return rc;
    CU(ctx, cudaMemcpyAsync(ok_host, dK, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    return DEX_OK;
}
static int host_eval_wait(dex_ctx* ctx) {
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
    return DEX_OK;
}

int dex_eval_host(dex_ctx* ctx, const dex_population* pop, const void* X_host, int32_t nfeatures,
                  int64_t nsamples, int64_t ldx, void* out_host, int64_t ldo, uint8_t* ok_host,
                  int eval_flags) {
    DEX_RANGE("dex_eval_host");
    int rc = host_eval_enqueue(ctx, pop, X_host, nfeatures, nsamples, ldx, out_host, ldo, ok_host, eval_flags);
    if (rc) return rc;
    return host_eval_wait(ctx);
}

// Sample-sharded evaluation over several devices from ONE host thread: device d
// evaluates the contiguous column block [N d / R, N (d+1) / R) of X against its own copy of the
// population and its rows land in place in the caller's (n_trees x nsamples) result.  Everything
// is enqueued on every device before anything is waited for, so the devices run concurrently; no
// collective is involved (the gather IS the strided device-to-host copy).
int dex_shard_eval_host(dex_ctx* const* ctxs, const dex_population* const* pops, int32_t n_devices,
                        const void* X_host, int32_t nfeatures, int64_t nsamples, int64_t ldx, void* out_host,
                        int64_t ldo, uint8_t* ok_host, int eval_flags) {
    DEX_RANGE("dex_shard_eval_host");
    if (!ctxs || !pops || n_devices < 1) return DEX_ERR_INVALID;
    for (int d = 0; d < n_devices; ++d)
        if (!ctxs[d] || !pops[d]) return DEX_ERR_INVALID;
    dex_ctx* c0 = ctxs[0];
    const int64_t P = pops[0]->h.n_trees;
    const size_t es = pops[0]->h.dtype == DEX_F32 ? 4 : 8;
    for (int d = 1; d < n_devices; ++d)
        if (pops[d]->h.n_trees != P || pops[d]->h.dtype != pops[0]->h.dtype)
            return set_err(c0, DEX_ERR_INVALID, "the per-device populations differ (pack the same trees on every device)");
    if (ldo < nsamples) return set_err(c0, DEX_ERR_INVALID, "ldo < nsamples");
    std::vector<uint8_t> flags((size_t)n_devices * (size_t)std::max<int64_t>(P, 1), 1);
    int rc = DEX_OK;
    int issued = 0;
    for (int d = 0; d < n_devices && rc == DEX_OK; ++d) {
        const int64_t s = nsamples * d / n_devices, e = nsamples * (d + 1) / n_devices;
        // (SKIP_INCOMPLETE would make the enqueue wait for each device's flags in turn: not here)
        rc = host_eval_enqueue(ctxs[d], pops[d], static_cast<const char*>(X_host) + (size_t)s * (size_t)ldx * es, nfeatures,
                               e - s, ldx, static_cast<char*>(out_host) + (size_t)s * es, ldo,
                               flags.data() + (size_t)d * (size_t)std::max<int64_t>(P, 1),
                               eval_flags & ~DEX_EVAL_SKIP_INCOMPLETE);
        if (rc != DEX_OK && ctxs[d] != c0) set_err(c0, rc, std::string("device ") + std::to_string(d) + ": " + ctxs[d]->last_error);
        ++issued;
    }
    for (int d = 0; d < issued; ++d) {
        const int rw = host_eval_wait(ctxs[d]);
        if (rc == DEX_OK && rw != DEX_OK) rc = rw;
    }
    if (rc != DEX_OK) return rc;
    // a tree is complete iff it is complete on every shard
    for (int64_t t = 0; t < P; ++t) {
        uint8_t k = 1;
        for (int d = 0; d < n_devices; ++d) k &= flags[(size_t)d * (size_t)P + (size_t)t];
        ok_host[t] = k;
    }
    return DEX_OK;
}

// ok[t] = AND over the R shards of flags[r * P + t]
__global__ void and_flags_kernel(const uint8_t* __restrict__ flags, int R, int64_t P, uint8_t* __restrict__ ok) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P) return;
    uint8_t k = 1;
    for (int r = 0; r < R; ++r) k &= flags[(size_t)r * (size_t)P + (size_t)t];
    ok[t] = k;
}

// Sample-sharded evaluation with the result gathered on ONE device of the same process (SURVEY.md
// §8b `dex_shard_eval` + `dex_gather` as one call): device d evaluates ITS column block (X_devs[d],
// resident on device d) and its interpreter kernel stores straight into the root device's
// (n_trees x nsamples) matrix through peer memory — the gather is the kernel's own streaming stores
// over NVLink, overlapped with the arithmetic, no staging and no collective.  Flags are AND-reduced on
// the root.  Asynchronous: everything is ordered behind the work already enqueued on ctxs[root]'s
// stream, and the results are complete in that stream's order (dex_ctx_synchronize(ctxs[root])).
int dex_shard_eval(dex_ctx* const* ctxs, const dex_population* const* pops, int32_t n_devices,
                   const void* const* X_devs, int32_t nfeatures, int64_t nsamples, int64_t ldx,
                   void* out_root_dev, int64_t ldo, uint8_t* ok_root_dev, int32_t root, int eval_flags) {
    DEX_RANGE("dex_shard_eval");
    if (!ctxs || !pops || !X_devs || n_devices < 1 || root < 0 || root >= n_devices) return DEX_ERR_INVALID;
    for (int d = 0; d < n_devices; ++d)
        if (!ctxs[d] || !pops[d]) return DEX_ERR_INVALID;
    dex_ctx* cr = ctxs[root];
    const int64_t P = pops[0]->h.n_trees;
    const size_t es = pops[0]->h.dtype == DEX_F32 ? 4 : 8;
    for (int d = 1; d < n_devices; ++d)
        if (pops[d]->h.n_trees != P || pops[d]->h.dtype != pops[0]->h.dtype)
            return set_err(cr, DEX_ERR_INVALID, "the per-device populations differ (pack the same trees on every device)");
    if (ldo < nsamples) return set_err(cr, DEX_ERR_INVALID, "ldo < nsamples");
    if (P == 0) return DEX_OK;
    if (!ok_root_dev || (!out_root_dev && nsamples > 0)) return set_err(cr, DEX_ERR_INVALID, "null out / ok");
    int rc = ensure_device(cr);
    if (rc) return rc;
    // the shards' flags land side by side in the root's staging buffer
    if ((rc = ensure_dev_io(cr, (size_t)n_devices * (size_t)P))) return rc;
    uint8_t* flags = static_cast<uint8_t*>(cr->dev_io);
    // peers run behind whatever the root stream holds so far (its output buffer may still be read)
    CU(cr, cudaEventRecord(cr->ev[0], cr->stream));
    for (int d = 0; d < n_devices; ++d) {
        dex_ctx* c = ctxs[d];
        if ((rc = ensure_device(c))) return rc;      // makes c->device current
        if (c->device != cr->device) {
            int can = 0;
            CU(c, cudaDeviceCanAccessPeer(&can, c->device, cr->device));
            if (!can)
                return set_err(cr, DEX_ERR_UNSUPPORTED, "device " + std::to_string(c->device) + " cannot access device " +
                                                             std::to_string(cr->device) + " (no peer path)");
            cudaError_t e = cudaDeviceEnablePeerAccess(cr->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return cuda_err(c, e, "cudaDeviceEnablePeerAccess");
        }
        if (c != cr) CU(c, cudaStreamWaitEvent(c->stream, cr->ev[0], 0));
        const int64_t s = nsamples * d / n_devices, e = nsamples * (d + 1) / n_devices;
        rc = dex_eval(c, pops[d], X_devs[d], nfeatures, e - s, ldx,
                      out_root_dev ? static_cast<char*>(out_root_dev) + (size_t)s * es : nullptr, ldo,
                      flags + (size_t)d * (size_t)P, eval_flags & ~DEX_EVAL_SKIP_INCOMPLETE);
        if (rc != DEX_OK) {
            if (c != cr) set_err(cr, rc, std::string("device ") + std::to_string(d) + ": " + c->last_error);
            return rc;
        }
    }
    if ((rc = ensure_device(cr))) return rc;
    for (int d = 0; d < n_devices; ++d)
        if (ctxs[d] != cr && ctxs[d]->ev_done_valid) CU(cr, cudaStreamWaitEvent(cr->stream, ctxs[d]->ev_done, 0));
    and_flags_kernel<<<(unsigned)((P + 255) / 256), 256, 0, cr->stream>>>(flags, n_devices, P, ok_root_dev);
    cr->launches += 1;
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return cuda_err(cr, le, "flag reduction");
    return finish_call(cr, DEX_OK);
}

// plain copies on the context's stream, for hosts that have no CUDA binding of their own (the
// Julia extension holds device buffers as raw pointers from dex_device_alloc)
int dex_copy_to_device(dex_ctx* ctx, void* dst_dev, const void* src_host, int64_t bytes) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (bytes < 0 || (bytes > 0 && (!dst_dev || !src_host))) return set_err(ctx, DEX_ERR_INVALID, "null pointer / negative size");
    if (bytes) CU(ctx, cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
    return DEX_OK;
}
// returns after the bytes have landed (synchronises the context's stream)
int dex_copy_to_host(dex_ctx* ctx, void* dst_host, const void* src_dev, int64_t bytes) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (bytes < 0 || (bytes > 0 && (!dst_host || !src_dev))) return set_err(ctx, DEX_ERR_INVALID, "null pointer / negative size");
    if (bytes) CU(ctx, cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return DEX_OK;
}

// pinned host buffers for callers of the *_host entry points
int dex_host_alloc(void** out, int64_t bytes) {
    if (!out || bytes < 0) return DEX_ERR_INVALID;
    if (cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 1)) != cudaSuccess) { cudaGetLastError(); return DEX_ERR_NOMEM; }
    return DEX_OK;
}
int dex_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); return DEX_ERR_CUDA; }
    return DEX_OK;
}

// ---- peer memory (CUDA IPC): see include/dexb200.h ------------------------------------------
int dex_device_alloc(dex_ctx* ctx, void** out, int64_t bytes) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (!out || bytes < 0) return set_err(ctx, DEX_ERR_INVALID, "null out / negative size");
    *out = nullptr;
    cudaError_t e = cudaMalloc(out, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(ctx, DEX_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    return DEX_OK;
}
int dex_device_free(dex_ctx* ctx, void* p) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (p) CU(ctx, cudaFree(p));
    return DEX_OK;
}
int dex_ipc_export(dex_ctx* ctx, const void* dev_ptr, uint8_t* handle64) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (!dev_ptr || !handle64) return set_err(ctx, DEX_ERR_INVALID, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == DEX_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle64, &h, sizeof(h));
    return DEX_OK;
}
int dex_ipc_open(dex_ctx* ctx, const uint8_t* handle64, void** out) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (!handle64 || !out) return set_err(ctx, DEX_ERR_INVALID, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    *out = nullptr;
    CU(ctx, cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return DEX_OK;
}
int dex_ipc_close(dex_ctx* ctx, void* p) {
    int rc = ensure_device(ctx);
    if (rc) return rc;
    if (p) CU(ctx, cudaIpcCloseMemHandle(p));
    return DEX_OK;
}

}  // extern "C"
