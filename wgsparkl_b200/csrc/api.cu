// C ABI (include/b200mpm.h) and host-side orchestration of one MPM substep.
// Mirrors MpmPipeline / MpmData (src/pipeline.rs:24-39, 84-95) — see INTEGRATION.md for the
// Rust-side binding. No CPU fallback: every entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "launch.h"
#include "nccl_dyn.h"

using namespace b2;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) {                                                                       \
            int _code = (_e == cudaErrorMemoryAllocation) ? B200MPM_ERR_OUT_OF_MEMORY : B200MPM_ERR_CUDA; \
            return fail(_code, std::string(#expr) + ": " + cudaGetErrorString(_e));                    \
        }                                                                                              \
    } while (0)

struct EventPair {
    cudaEvent_t a, b;
    int kernel, pass; // B200MPM_KERNEL_*, and the reference pass (B200MPM_PASS_*) the kernel belongs to
};

} // namespace

struct b200mpm_pipeline {
    int device = 0;
    int dim = 3;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    uint64_t launches = 0;
    bool timestamps = false;
    int cur_pass = 0; // the reference pass the kernel timers are booked on (PassTimer)
    bool use_graphs = true;
    std::vector<EventPair> events;
    std::vector<cudaEvent_t> event_pool;
    double pass_ms[B200MPM_NUM_PASSES] = {0};
    double kernel_ms[B200MPM_NUM_KERNELS] = {0};
    std::vector<b200mpm_data*> children; // data objects created on this pipeline (orphaned, not freed, on destroy)
    // scratch for b200mpm_prefix_sum_u32
    LaunchCfg cfg() { return LaunchCfg{dim, num_sms, stream, &launches}; }
};

struct b200mpm_data {
    b200mpm_pipeline* pipe = nullptr; // null once the pipeline has been destroyed (only destroy is legal then)
    int device = 0;
    DeviceData dev{};
    int cur = 0;
    bool sorted_indirect = true; // sorted_ids is an indirection into `cur` (no full substep since the last sort)
    // The hash map / per-cell bins still hold a sort (after creation, a grid reallocation or b200mpm_sort_only): the
    // next substep has to launch k_begin_substep itself. A complete substep leaves the grid clean (the tail of its
    // k_g2p is the clearing).
    bool grid_dirty = true;
    uint32_t num_bodies = 0;
    std::vector<void*> allocs;
    std::vector<void*> grid_allocs; // the capacity-sized arrays (replaced by b200mpm_data_reserve_grid)
    float auto_grow_load = 0.0f; // > 0: grow the grid when more than this fraction of the block capacity is active
    uint32_t particle_cap = 0;
    void* staging = nullptr; // device staging for readbacks / host writes
    size_t staging_bytes = 0;
    void* pinned = nullptr; // pinned host mirror of small transfers
    size_t pinned_bytes = 0;
    // One captured CUDA graph per ping-pong parity: the substep is recorded once and replayed, like the
    // reference's KernelInvocationQueue (src_testbed/step.rs:122-128). Independent kernels sit on parallel
    // branches of the graph.
    cudaGraphExec_t graph_exec[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}}; // [parity][phase]
    uint64_t graph_launches[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
    // Native slab exchange (b200mpm_shard_comm_init): NCCL communicator + device exchange buffers.
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    void* mig_send[2] = {nullptr, nullptr};
    void* mig_recv[2] = {nullptr, nullptr};
    void* halo_send[2] = {nullptr, nullptr};
    void* halo_recv[2] = {nullptr, nullptr};
    int* imp_buf = nullptr;
    uint32_t mig_cap = 0, halo_cap = 0;
    // Peer-to-peer exchange (b200mpm_shard_p2p_export / _connect): one arena per rank, mapped by both neighbours.
    //   arena = [flags: 4 x u32 (+pad to 256 B)] [mig from-left x2 parity] [mig from-right x2] [halo from-left x2] [halo from-right x2]
    bool p2p_ready = false;
    char* arena = nullptr;
    char* peer_arena[2] = {nullptr, nullptr}; // -x / +x neighbour's arena (IPC mapping)
    size_t arena_mig_bytes = 0, arena_halo_bytes = 0;
    bool bodies_react = false; // some body can react to an impulse (mass or motion): the impulse all-reduce is needed
    uint32_t n_live_host = 0; // host mirror of counters->n_live (sharded runs track it)
    bool sharded = false;
    // The ORDERED read-backs scatter by original particle id into a buffer of num_particles entries: only valid when
    // the ids are the default 0..n-1 and there is no spare capacity (b200mpm_data_create_ex may break both).
    bool ordered_readback_ok = true;
    // Asynchronous position readback (b200mpm_read_positions_async): two device staging slots, a copy stream.
    cudaStream_t copy_stream = nullptr;
    float4* pos_stage[2] = {nullptr, nullptr};
    cudaEvent_t pos_gathered[2] = {nullptr, nullptr}, pos_copied[2] = {nullptr, nullptr};
    int pos_slot = 0;
};

namespace {

void orphan_data(b200mpm_data* d) { d->pipe = nullptr; }

template <class T>
int dev_alloc(b200mpm_data* d, T** out, size_t count, bool zero = true, std::vector<void*>* list = nullptr) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CU_TRY(cudaMalloc(&p, bytes));
    (list ? list : &d->allocs)->push_back(p);
    if (zero) CU_TRY(cudaMemsetAsync(p, 0, bytes, d->pipe->stream));
    *out = (T*)p;
    return 0;
}

int ensure_staging(b200mpm_data* d, size_t bytes) {
    if (d->staging_bytes >= bytes) return 0;
    if (d->staging) cudaFree(d->staging);
    d->staging = nullptr;
    d->staging_bytes = 0;
    CU_TRY(cudaMalloc(&d->staging, bytes));
    d->staging_bytes = bytes;
    return 0;
}
int ensure_pinned(b200mpm_data* d, size_t bytes) {
    if (d->pinned_bytes >= bytes) return 0;
    if (d->pinned) cudaFreeHost(d->pinned);
    d->pinned = nullptr;
    d->pinned_bytes = 0;
    CU_TRY(cudaMallocHost(&d->pinned, bytes));
    d->pinned_bytes = bytes;
    return 0;
}

// The sparse grid's arrays, all sized by the block capacity (= hash capacity, a power of two; grid.rs:283).
// Their contents are rebuilt by every substep, so a reallocation does not have to carry anything over.
int alloc_grid(b200mpm_data* d, uint32_t capacity) {
    DeviceData& dev = d->dev;
    const int na = (d->pipe->dim == 2) ? 4 : 8;
    auto* L = &d->grid_allocs;
    int r = 0;
    if ((r = dev_alloc(d, &dev.hkeys, capacity, true, L))) return r;
    if ((r = dev_alloc(d, &dev.hvals, capacity, true, L))) return r;
    if ((r = dev_alloc(d, &dev.block_vid, capacity, true, L))) return r;
    if ((r = dev_alloc(d, &dev.cell_start, (size_t)capacity * CELLS_PER_BLOCK + 1, true, L))) return r;
    if ((r = dev_alloc(d, &dev.nbr, (size_t)capacity * na, true, L))) return r;
    if ((r = dev_alloc(d, &dev.node_mv, (size_t)capacity * CELLS_PER_BLOCK, true, L))) return r;
    if (dev.has_bodies && (r = dev_alloc(d, &dev.node_cdf, (size_t)capacity * CELLS_PER_BLOCK, true, L))) return r;
    if (dev.has_bodies && (r = dev_alloc(d, &dev.node_imp, (size_t)capacity * CELLS_PER_BLOCK * 2, true, L))) return r;
    if ((r = dev_alloc(d, &dev.block_flags, capacity, true, L))) return r;
    if ((r = dev_alloc(d, &dev.block_f0, capacity, true, L))) return r;
    if ((r = dev_alloc(d, &dev.cpic_list, capacity, true, L))) return r;
    dev.g2p_items_len = capacity + d->particle_cap / G2P_ITEM + 1;
    if ((r = dev_alloc(d, &dev.g2p_items, dev.g2p_items_len, true, L))) return r;
    if ((r = dev_alloc(d, &dev.p2g_list, (size_t)P2G_BUCKETS * capacity, true, L))) return r;
    if ((r = dev_alloc(d, &dev.block_range, capacity, true, L))) return r;
    if ((r = dev_alloc(d, &dev.scan_state, scan_num_tiles((uint64_t)capacity * CELLS_PER_BLOCK + 1) + 2, true, L))) return r;
    dev.capacity = capacity;
    return 0;
}

// queue.compute_pass(name, add_timestamps) (src/pipeline.rs:201). ONE pair of events per kernel: a PassTimer with a
// pass id only names the reference pass that the kernel timers inside it are booked on (a second, nested pair of
// events per pass added ~6 us of event overhead to every kernel measured inside it).
struct PassTimer {
    b200mpm_pipeline* p;
    cudaEvent_t a = nullptr;
    int id, outer_pass = 0; // pass id, or B200MPM_NUM_PASSES + kernel id for the kernel-level timers
    PassTimer(b200mpm_pipeline* pipe, int pass_or_kernel) : p(pipe), id(pass_or_kernel) {
        if (!p->timestamps) return;
        if (id < B200MPM_NUM_PASSES) {
            outer_pass = p->cur_pass;
            p->cur_pass = id;
            return;
        }
        a = take();
        cudaEventRecord(a, p->stream);
    }
    cudaEvent_t take() {
        if (!p->event_pool.empty()) {
            cudaEvent_t e = p->event_pool.back();
            p->event_pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    ~PassTimer() {
        if (!p->timestamps) return;
        if (id < B200MPM_NUM_PASSES) {
            p->cur_pass = outer_pass;
            return;
        }
        cudaEvent_t b = take();
        cudaEventRecord(b, p->stream);
        p->events.push_back(EventPair{a, b, id - B200MPM_NUM_PASSES, p->cur_pass});
    }
};

void fold_events(b200mpm_pipeline* p) {
    if (p->events.empty()) return;
    cudaStreamSynchronize(p->stream);
    for (auto& e : p->events) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        p->pass_ms[e.pass] += ms;
        p->kernel_ms[e.kernel] += ms;
        p->event_pool.push_back(e.a);
        p->event_pool.push_back(e.b);
    }
    p->events.clear();
}

void run_sort(b200mpm_pipeline* p, b200mpm_data* d) {
    LaunchCfg c = p->cfg();
    constexpr int K = B200MPM_NUM_PASSES; // kernel-level timers nest inside the reference's pass timers
    {
        PassTimer t(p, B200MPM_PASS_GRID_SORT);
        {
            PassTimer k(p, K + B200MPM_KERNEL_TOUCH);
            launch_touch(c, d->dev, d->cur);
        }
        {
            PassTimer k(p, K + B200MPM_KERNEL_RIGID);
            launch_touch_rigid(c, d->dev);
        }
    }
    {
        PassTimer t(p, B200MPM_PASS_GRID_UPDATE_CDF);
        PassTimer k(p, K + B200MPM_KERNEL_BLOCK_PREPARE);
        launch_block_prepare(c, d->dev);
    }
    {
        PassTimer t(p, B200MPM_PASS_P2G_CDF);
        PassTimer k(p, K + B200MPM_KERNEL_RIGID);
        launch_p2g_cdf(c, d->dev);
    }
    {
        PassTimer t(p, B200MPM_PASS_GRID_SORT);
        PassTimer k(p, K + B200MPM_KERNEL_SCATTER);
        launch_scatter(c, d->dev, d->cur);
    }
}

// Enqueues one substep, a plain chain of five kernels (seven launches more with mesh colliders, [..]):
//   [transform_rigid ->] touch (+ count, + the previous substep's integrate) [-> mark_rigid -> touch_rigid]
//     -> block_prepare (+ the blocks' ranges) [-> p2g_cdf] -> scatter -> p2g -> g2p (+ the clearing for the next substep)
//     [-> integrate]
// Round 1 forked the independent kernels onto a second stream; the persistent kernels take every SM slot, so the
// side branch never ran beside them, only before or after (profiles/r02_timeline.md) - each was folded into the
// kernel it used to run next to.
// `deferrable`: the caller replays this substep back to back (graph capture), see defer_integrate.
// phase: PHASE_ALL = the whole substep; PHASE_BEGIN = up to and including P2G; PHASE_END = from G2P on
// (sharded runs exchange the node halo and the body impulses between the two).
enum { PHASE_ALL = 0, PHASE_BEGIN = 1, PHASE_END = 2, PHASE_SHARDED = 3 };

void enqueue_substep(b200mpm_pipeline* p, b200mpm_data* d, cudaStream_t main, bool deferrable, uint64_t* counter,
                     int phase, bool outer_defers = false) {
    LaunchCfg c{p->dim, p->num_sms, main, counter};
    const DeviceData& dev = d->dev;
    // The body integration (<= 16 bodies, one warp: a pure latency chain of ~7 us) is DEFERRED in the single-GPU
    // graph: it rides in an extra CTA of the next substep's k_touch, i.e. in front of k_block_prepare (the first kernel
    // that looks at the poses again); b200mpm_step flushes the last one.
    // (mesh colliders: k_transform_rigid needs the poses first; sharded: the peer-to-peer substep only - its PHASE_BEGIN
    // half is told through outer_defers)
    const bool defer_integrate = outer_defers || (deferrable && dev.num_rigid == 0 &&
                                                  (phase == PHASE_ALL || (phase == PHASE_SHARDED && d->p2p_ready)));
    auto finish_substep = [&]() {
        launch_g2p_update(c, dev, d->cur); // (its retiring CTAs clear the hash map and the bins for the next substep)
        if (!defer_integrate) launch_integrate_bodies(c, dev);
    };
    if (phase == PHASE_SHARDED) {
        // One whole substep of a slab: migration and node-halo exchanges are NCCL send/recv groups on the same
        // stream (and inside the same captured graph) as the kernels.
        const NcclApi& nc = nccl_api();
        const int left = d->rank > 0 ? d->rank - 1 : -1, right = d->rank < d->world - 1 ? d->rank + 1 : -1;
        const size_t mig_bytes = B200MPM_SHARD_HEADER_BYTES + (size_t)d->mig_cap * B200MPM_PARTICLE_RECORD_BYTES;
        const size_t halo_bytes = B200MPM_SHARD_HEADER_BYTES + (size_t)d->halo_cap * B200MPM_HALO_BLOCK_BYTES;
        auto exchange = [&](void** send, void** recv, size_t bytes) {
            nc.GroupStart();
            if (left >= 0) {
                nc.Send(send[0], bytes, ncclUint8, left, d->comm, main);
                nc.Recv(recv[0], bytes, ncclUint8, left, d->comm, main);
            }
            if (right >= 0) {
                nc.Send(send[1], bytes, ncclUint8, right, d->comm, main);
                nc.Recv(recv[1], bytes, ncclUint8, right, d->comm, main);
            }
            nc.GroupEnd();
        };
        if (d->p2p_ready) {
            // Peer-to-peer: the pack kernels store straight into the neighbour's receive buffers over NVLink.
            const int par = d->cur; // double buffering by substep parity
            auto mig_buf = [&](char* arena, int from_dir) { return arena + 256 + (size_t)(from_dir * 2 + par) * d->arena_mig_bytes; };
            auto halo_buf = [&](char* arena, int from_dir) {
                return arena + 256 + 4 * d->arena_mig_bytes + (size_t)(from_dir * 2 + par) * d->arena_halo_bytes;
            };
            auto flag = [&](char* arena, int from_dir, int kind) { return (uint32_t*)arena + from_dir * 2 + kind; };
            char* pl = d->peer_arena[0];
            char* pr = d->peer_arena[1];
            // I am the +x neighbour of `left` (its from-right buffers) and the -x neighbour of `right`.
            void* ml = pl ? (void*)mig_buf(pl, 1) : d->mig_send[0];
            void* mr = pr ? (void*)mig_buf(pr, 0) : d->mig_send[1];
            void* hl = pl ? (void*)halo_buf(pl, 1) : d->halo_send[0];
            void* hr = pr ? (void*)halo_buf(pr, 0) : d->halo_send[1];
            // tick + the emigrants k_g2p listed + publication: one CTA
            launch_emigrate_listed(c, dev, d->cur, ml, mr, d->mig_cap, pl ? flag(pl, 1, 0) : nullptr, pr ? flag(pr, 0, 0) : nullptr);
            launch_immigrate_p2p(c, dev, d->cur, pl ? mig_buf(d->arena, 0) : nullptr, pr ? mig_buf(d->arena, 1) : nullptr,
                                 flag(d->arena, 0, 0), flag(d->arena, 1, 0), d->mig_cap);
            enqueue_substep(p, d, main, deferrable, counter, PHASE_BEGIN, defer_integrate);
            launch_halo_pack(c, dev, hl, hr, d->halo_cap, pl ? flag(pl, 1, 1) : nullptr, pr ? flag(pr, 0, 1) : nullptr, true);
            launch_halo_add(c, dev, left >= 0 ? halo_buf(d->arena, 0) : nullptr, right >= 0 ? halo_buf(d->arena, 1) : nullptr,
                            d->halo_cap, flag(d->arena, 0, 1), flag(d->arena, 1, 1));
            if (d->bodies_react) {
                launch_impulses_io(c, dev, d->imp_buf, 0);
                nc.AllReduce(d->imp_buf, d->imp_buf, B200MPM_MAX_BODIES * 6, ncclInt32, ncclSum, d->comm, main); // exact
                launch_impulses_io(c, dev, d->imp_buf, 1);
            }
            // (the dead tail of this substep is dropped by the next k_shard_tick)
            finish_substep();
            return;
        } else {
            launch_emigrate(c, dev, d->cur, d->mig_send[0], d->mig_send[1], d->mig_cap);
            exchange(d->mig_send, d->mig_recv, mig_bytes);
            if (left >= 0) launch_immigrate(c, dev, d->cur, d->mig_recv[0], d->mig_cap);
            if (right >= 0) launch_immigrate(c, dev, d->cur, d->mig_recv[1], d->mig_cap);
            enqueue_substep(p, d, main, deferrable, counter, PHASE_BEGIN);
            launch_halo_pack(c, dev, d->halo_send[0], d->halo_send[1], d->halo_cap);
            exchange(d->halo_send, d->halo_recv, halo_bytes);
            launch_halo_add(c, dev, left >= 0 ? d->halo_recv[0] : nullptr, right >= 0 ? d->halo_recv[1] : nullptr, d->halo_cap);
        }
        if (d->bodies_react) {
            launch_impulses_io(c, dev, d->imp_buf, 0);
            nc.AllReduce(d->imp_buf, d->imp_buf, B200MPM_MAX_BODIES * 6, ncclInt32, ncclSum, d->comm, main); // exact
            launch_impulses_io(c, dev, d->imp_buf, 1);
        }
        enqueue_substep(p, d, main, deferrable, counter, PHASE_END);
        return;
    }
    if (phase == PHASE_END) {
        finish_substep();
        launch_drop_dead_tail(c, dev);
        return;
    }
    launch_transform_rigid(c, dev);
    launch_touch(c, dev, d->cur, defer_integrate); // (+ the previous substep's body integration, no-op on the first)
    launch_touch_rigid(c, dev);
    launch_block_prepare(c, dev); // (+ the blocks' ranges in the sorted array: what used to be the scan)
    launch_p2g_cdf(c, dev);
    launch_scatter(c, dev, d->cur);
    launch_p2g(c, dev, d->cur); // (with bodies: particle colouring "g2p_cdf" + collider-side blocks + all others, one kernel)
    if (phase == PHASE_BEGIN) return;
    finish_substep();
}

// The sparse grid must be clear before a sort starts (see b200mpm_data::grid_dirty).
void ensure_clean_grid(b200mpm_pipeline* p, b200mpm_data* d) {
    if (!d->grid_dirty) return;
    launch_begin_substep(p->cfg(), d->dev);
    d->grid_dirty = false;
}

// Captures the substep for the current parity into a graph (once), then replays it.
bool run_substep_graph(b200mpm_pipeline* p, b200mpm_data* d, int phase) {
    const int par = d->cur;
    if (!d->graph_exec[par][phase]) {
        cudaGraph_t graph = nullptr;
        uint64_t count = 0;
        if (cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        enqueue_substep(p, d, p->stream, true, &count, phase);
        if (cudaStreamEndCapture(p->stream, &graph) != cudaSuccess || !graph) {
            cudaGetLastError();
            return false;
        }
        cudaError_t e = cudaGraphInstantiate(&d->graph_exec[par][phase], graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            cudaGetLastError();
            d->graph_exec[par][phase] = nullptr;
            return false;
        }
        d->graph_launches[par][phase] = count;
    }
    if (phase != PHASE_END) ensure_clean_grid(p, d);
    if (cudaGraphLaunch(d->graph_exec[par][phase], p->stream) != cudaSuccess) return false;
    p->launches += d->graph_launches[par][phase];
    d->grid_dirty = (phase == PHASE_BEGIN); // every other phase ends with the clearing for the next substep
    if (phase != PHASE_BEGIN) {
        d->cur ^= 1;
        d->sorted_indirect = false;
    }
    return true;
}

// One phase of a substep without timestamps: graph replay, or plain launches when graphs are disabled.
void run_phase(b200mpm_pipeline* p, b200mpm_data* d, int phase) {
    if (p->use_graphs && run_substep_graph(p, d, phase)) return;
    if (phase != PHASE_END) ensure_clean_grid(p, d);
    enqueue_substep(p, d, p->stream, false, &p->launches, phase);
    d->grid_dirty = (phase == PHASE_BEGIN);
    if (phase != PHASE_BEGIN) {
        d->cur ^= 1;
        d->sorted_indirect = false;
    }
}

void run_substep(b200mpm_pipeline* p, b200mpm_data* d) {
    if (!p->timestamps && p->use_graphs && run_substep_graph(p, d, PHASE_ALL)) return;
    LaunchCfg c = p->cfg();
    {
        PassTimer t(p, B200MPM_PASS_UPDATE_RIGID_PARTICLES);
        PassTimer k(p, B200MPM_NUM_PASSES + B200MPM_KERNEL_RIGID);
        ensure_clean_grid(p, d);
        launch_transform_rigid(c, d->dev);
    }
    run_sort(p, d);
    constexpr int K = B200MPM_NUM_PASSES;
    {
        // (the "g2p_cdf" pass runs inside the P2G kernel, on the particles of collider-side blocks while they are staged)
        PassTimer t(p, B200MPM_PASS_P2G);
        PassTimer k(p, K + B200MPM_KERNEL_P2G);
        launch_p2g(c, d->dev, d->cur);
    }
    {
        // grid_update + g2p + particles_update are one kernel here (whose tail is reset_hmap for the next substep)
        PassTimer t(p, B200MPM_PASS_G2P);
        PassTimer k(p, K + B200MPM_KERNEL_G2P);
        launch_g2p_update(c, d->dev, d->cur);
    }
    {
        PassTimer t(p, B200MPM_PASS_INTEGRATE_BODIES);
        PassTimer k(p, K + B200MPM_KERNEL_INTEGRATE_BODIES);
        launch_integrate_bodies(c, d->dev);
    }
    d->cur ^= 1;
    d->sorted_indirect = false;
    if (p->timestamps && p->events.size() > 4096) fold_events(p);
}

} // namespace

extern "C" {

const char* b200mpm_last_error(void) { return g_last_error.c_str(); }

int b200mpm_pipeline_create(int device, int dim, b200mpm_pipeline** out) {
    if (!out) return fail(B200MPM_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (dim != 2 && dim != 3) return fail(B200MPM_ERR_INVALID_ARGUMENT, "dim must be 2 or 3");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(B200MPM_ERR_NO_DEVICE, "no CUDA device is visible; this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return fail(B200MPM_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(B200MPM_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100; kernels are built for sm_100a only");
    CU_TRY(cudaSetDevice(device));
    auto* p = new b200mpm_pipeline();
    p->device = device;
    p->dim = dim;
    p->num_sms = prop.multiProcessorCount;
    p->use_graphs = getenv("B200MPM_NO_GRAPH") == nullptr; // debugging aid: plain launches
    cudaError_t e = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete p;
        return fail(B200MPM_ERR_CUDA, cudaGetErrorString(e));
    }
    p->stream = p->own_stream;
    *out = p;
    return B200MPM_OK;
}

void b200mpm_pipeline_destroy(b200mpm_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    for (auto& e : p->events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto e : p->event_pool) cudaEventDestroy(e);
    // Data objects that outlive their pipeline stay valid for b200mpm_data_destroy only.
    for (b200mpm_data* d : p->children) orphan_data(d);
    cudaStreamDestroy(p->own_stream);
    delete p;
}

int b200mpm_pipeline_set_stream(b200mpm_pipeline* p, void* cuda_stream) {
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null pipeline");
    CU_TRY(cudaSetDevice(p->device));
    fold_events(p);
    CU_TRY(cudaStreamSynchronize(p->stream));
    p->stream = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
    return B200MPM_OK;
}

uint64_t b200mpm_pipeline_launch_count(const b200mpm_pipeline* p) { return p ? p->launches : 0; }

int b200mpm_data_create(b200mpm_pipeline* p, const b200mpm_sim_params* params, const b200mpm_particle* particles,
                        size_t num_particles, const b200mpm_body* bodies, size_t num_bodies, float cell_width,
                        uint32_t grid_capacity, b200mpm_data** out) {
    return b200mpm_data_create_ex(p, params, particles, num_particles, nullptr, num_particles, bodies, num_bodies, cell_width,
                                  grid_capacity, out);
}

int b200mpm_data_create_ex(b200mpm_pipeline* p, const b200mpm_sim_params* params, const b200mpm_particle* particles,
                           size_t num_particles, const uint32_t* particle_ids, size_t particle_capacity,
                           const b200mpm_body* bodies, size_t num_bodies, float cell_width, uint32_t grid_capacity,
                           b200mpm_data** out) {
    if (!p || !params || !out) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (particle_capacity < num_particles) return fail(B200MPM_ERR_INVALID_ARGUMENT, "particle_capacity < num_particles");
    if (particle_capacity >= (1ull << 31)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "too many particles");
    *out = nullptr;
    if (num_particles && !particles) return fail(B200MPM_ERR_INVALID_ARGUMENT, "particles is null");
    if (num_bodies && !bodies) return fail(B200MPM_ERR_INVALID_ARGUMENT, "bodies is null");
    if (num_bodies > B200MPM_MAX_BODIES)
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "at most 16 coupled colliders are supported (rigid_impulses.rs:42)");
    if (!(cell_width > 0.0f)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "cell_width must be positive");
    if (grid_capacity == 0 || grid_capacity > (1u << 24))
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "grid_capacity must be in [1, 2^24]");
    if (num_particles >= (1ull << 31)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "too many particles");
    for (size_t i = 0; i < num_bodies; ++i)
        if (bodies[i].shape_type > B200MPM_SHAPE_CAPSULE &&
            bodies[i].shape_type != (uint32_t)(p->dim == 3 ? B200MPM_SHAPE_TRIMESH : B200MPM_SHAPE_POLYLINE))
            return fail(B200MPM_ERR_INVALID_ARGUMENT, "unsupported collider shape (ball, cuboid, capsule; trimesh in 3D, polyline in 2D)");
    CU_TRY(cudaSetDevice(p->device));

    uint32_t capacity = 1;
    while (capacity < grid_capacity) capacity <<= 1; // grid.rs:283
    const int D = p->dim;
    const uint32_t n = (uint32_t)num_particles; // particles uploaded now
    const uint32_t ncap = (uint32_t)particle_capacity; // room for immigrants (sharded runs)

    auto* d = new b200mpm_data();
    d->pipe = p;
    d->device = p->device;
    d->num_bodies = (uint32_t)num_bodies;
    DeviceData& dev = d->dev;
    dev.n = ncap;
    dev.capacity = capacity;
    d->n_live_host = n;
    d->ordered_readback_ok = (particle_ids == nullptr) && (particle_capacity == num_particles);
    dev.has_bodies = num_bodies > 0;

    // ---- material table (dedup of the per-particle model buffers) + SoA staging on the host
    std::vector<Material> materials;
    std::unordered_map<std::string, uint32_t> mat_index;
    std::vector<float4> pos4(n), vel4(n), Fa(n), Ca(n), Fb, Cb, plastic;
    std::vector<float> Fc, Cc;
    std::vector<uint32_t> aff;
    std::vector<float4> cdf_nd, cdf_rv;
    if (D == 3) {
        Fb.resize(n);
        Cb.resize(n);
        Fc.resize(n);
        Cc.resize(n);
    }
    bool has_plastic = false;
    plastic.resize(n);
    if (dev.has_bodies) {
        aff.resize(n);
        cdf_nd.resize(n);
        cdf_rv.resize(n);
    }
    for (uint32_t i = 0; i < n; ++i) {
        const b200mpm_particle& q = particles[i];
        Material m{};
        m.mass = q.mass;
        m.init_volume = q.init_volume;
        m.init_radius = q.init_radius;
        m.lambda = q.lambda;
        m.mu = q.mu;
        m.dp_h0 = q.dp_h0, m.dp_h1 = q.dp_h1, m.dp_h2 = q.dp_h2, m.dp_h3 = q.dp_h3;
        m.dp_lambda = q.dp_lambda, m.dp_mu = q.dp_mu;
        m.phase = q.phase;
        m.max_stretch = q.max_stretch;
        m.model = q.model;
        std::string key((const char*)&m, sizeof(Material));
        auto it = mat_index.find(key);
        uint32_t mid;
        if (it == mat_index.end()) {
            mid = (uint32_t)materials.size();
            if (mid > MAT_ID_MASK) {
                delete d;
                return fail(B200MPM_ERR_INVALID_ARGUMENT, "too many distinct materials");
            }
            materials.push_back(m);
            materials.back().dp_ratio = ((float)D * m.dp_lambda + 2.0f * m.dp_mu) / (2.0f * m.dp_mu);
            mat_index.emplace(std::move(key), mid);
        } else {
            mid = it->second;
        }
        // A particle can reach phase == 0 (plastic) if it starts there or can break by stretch
        // (particle_update.wgsl:101-122).
        if (q.phase == 0.0f || (q.phase > 0.0f && q.max_stretch > 0.0f && q.max_stretch < 3.0e38f)) has_plastic = true;
        union {
            uint32_t u;
            float f;
        } mb, ob;
        mb.u = mid;
        ob.u = particle_ids ? particle_ids[i] : i;
        pos4[i] = make_float4(q.position[0], q.position[1], D == 3 ? q.position[2] : 0.0f, mb.f);
        vel4[i] = make_float4(q.velocity[0], q.velocity[1], D == 3 ? q.velocity[2] : 0.0f, ob.f);
        Fa[i] = make_float4(q.def_grad[0], q.def_grad[1], q.def_grad[2], q.def_grad[3]);
        Ca[i] = make_float4(q.affine[0], q.affine[1], q.affine[2], q.affine[3]);
        if (D == 3) {
            Fb[i] = make_float4(q.def_grad[4], q.def_grad[5], q.def_grad[6], q.def_grad[7]);
            Fc[i] = q.def_grad[8];
            Cb[i] = make_float4(q.affine[4], q.affine[5], q.affine[6], q.affine[7]);
            Cc[i] = q.affine[8];
        }
        plastic[i] = make_float4(q.plastic_det, q.plastic_hardening, q.plastic_log_vol_gain, 0.0f);
        if (dev.has_bodies) {
            aff[i] = q.cdf_affinity;
            cdf_nd[i] = make_float4(q.cdf_normal[0], q.cdf_normal[1], D == 3 ? q.cdf_normal[2] : 0.0f, q.cdf_signed_distance);
            cdf_rv[i] = make_float4(q.cdf_rigid_vel[0], q.cdf_rigid_vel[1], D == 3 ? q.cdf_rigid_vel[2] : 0.0f, 0.0f);
        }
    }
    dev.has_plastic = has_plastic;
    dev.num_materials = (uint32_t)materials.size();

#define ALLOC(ptr, count)                                     \
    do {                                                      \
        int _r = dev_alloc(d, &(ptr), (count));               \
        if (_r != 0) {                                        \
            b200mpm_data_destroy(d);                          \
            return _r;                                        \
        }                                                     \
    } while (0)
#define UPLOAD(ptr, vec)                                                                                          \
    do {                                                                                                          \
        if (!(vec).empty()) {                                                                                     \
            cudaError_t _e = cudaMemcpyAsync((ptr), (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice, p->stream); \
            if (_e != cudaSuccess) {                                                                              \
                b200mpm_data_destroy(d);                                                                          \
                return fail(B200MPM_ERR_CUDA, cudaGetErrorString(_e));                                            \
            }                                                                                                     \
        }                                                                                                         \
    } while (0)

    for (int s = 0; s < 2; ++s) {
        ALLOC(dev.pos4[s], ncap);
        ALLOC(dev.vel4[s], ncap);
        ALLOC(dev.Fa[s], ncap);
        ALLOC(dev.Ca[s], ncap);
        if (D == 3) {
            ALLOC(dev.Fb[s], ncap);
            ALLOC(dev.Fc[s], ncap);
            ALLOC(dev.Cb[s], ncap);
            ALLOC(dev.Cc[s], ncap);
        }
        if (has_plastic) ALLOC(dev.plastic[s], ncap);
        if (dev.has_bodies) ALLOC(dev.cdf_aff[s], ncap);
    }
    if (dev.has_bodies) {
        ALLOC(dev.cdf_nd, ncap);
        ALLOC(dev.cdf_rv, ncap);
    }
    Material* dmat = nullptr;
    ALLOC(dmat, materials.size());
    dev.materials = dmat;
    ALLOC(dev.pkey, ncap);
    ALLOC(dev.rank, ncap);
    ALLOC(dev.sorted_ids, ncap);
    d->particle_cap = ncap;
    {
        int _r = alloc_grid(d, capacity);
        if (_r != 0) {
            b200mpm_data_destroy(d);
            return _r;
        }
    }
    ALLOC(dev.bodies, B200MPM_MAX_BODIES);
    ALLOC(dev.sim, 1);
    ALLOC(dev.counters, 1);
    ALLOC(dev.timeline, 2 * B200MPM_NUM_KERNELS);

    UPLOAD(dev.pos4[0], pos4);
    UPLOAD(dev.vel4[0], vel4);
    UPLOAD(dev.Fa[0], Fa);
    UPLOAD(dev.Ca[0], Ca);
    if (D == 3) {
        UPLOAD(dev.Fb[0], Fb);
        UPLOAD(dev.Fc[0], Fc);
        UPLOAD(dev.Cb[0], Cb);
        UPLOAD(dev.Cc[0], Cc);
    }
    if (has_plastic) UPLOAD(dev.plastic[0], plastic);
    if (dev.has_bodies) {
        UPLOAD(dev.cdf_aff[0], aff);
        UPLOAD(dev.cdf_nd, cdf_nd);
        UPLOAD(dev.cdf_rv, cdf_rv);
    }
    UPLOAD(dmat, materials);

    for (size_t i = 0; i < num_bodies; ++i) {
        const b200mpm_body& s = bodies[i];
        for (int k = 0; k < 3; ++k)
            if (s.linvel[k] != 0.0f || s.angvel[k] != 0.0f || (s.two_ways && s.inv_mass[k] != 0.0f)) d->bodies_react = true;
        for (int k = 0; k < 9; ++k)
            if (s.two_ways && s.inv_inertia[k] != 0.0f) d->bodies_react = true;
    }
    dev.bodies_react = d->bodies_react ? 1 : 0;
    std::vector<BodyDev> hb(B200MPM_MAX_BODIES);
    std::memset(hb.data(), 0, hb.size() * sizeof(BodyDev));
    for (size_t i = 0; i < num_bodies; ++i) {
        const b200mpm_body& s = bodies[i];
        BodyDev& b = hb[i];
        b.shape_type = s.shape_type;
        for (int k = 0; k < 3; ++k) {
            b.shape_a[k] = s.shape_a[k];
            b.shape_b[k] = s.shape_b[k];
            b.trans[k] = s.translation[k];
            b.linvel[k] = s.linvel[k];
            b.angvel[k] = s.angvel[k];
            b.local_inv_mass[k] = s.two_ways ? s.inv_mass[k] : 0.0f;
            b.local_com[k] = s.local_com[k];
        }
        b.radius = s.radius;
        for (int k = 0; k < 4; ++k) b.rot_raw[k] = s.rotation[k];
        for (int k = 0; k < 9; ++k) b.local_inv_inertia[k] = s.two_ways ? s.inv_inertia[k] : 0.0f;
        if (D == 2) {
            b.rot[0] = s.rotation[0], b.rot[1] = s.rotation[1], b.rot[2] = -s.rotation[1], b.rot[3] = s.rotation[0];
        } else {
            float i_ = s.rotation[0], j_ = s.rotation[1], k_ = s.rotation[2], w_ = s.rotation[3];
            b.rot[0] = 1.0f - 2.0f * (j_ * j_ + k_ * k_);
            b.rot[1] = 2.0f * (i_ * j_ + k_ * w_);
            b.rot[2] = 2.0f * (i_ * k_ - j_ * w_);
            b.rot[3] = 2.0f * (i_ * j_ - k_ * w_);
            b.rot[4] = 1.0f - 2.0f * (i_ * i_ + k_ * k_);
            b.rot[5] = 2.0f * (j_ * k_ + i_ * w_);
            b.rot[6] = 2.0f * (i_ * k_ + j_ * w_);
            b.rot[7] = 2.0f * (j_ * k_ - i_ * w_);
            b.rot[8] = 1.0f - 2.0f * (i_ * i_ + j_ * j_);
        }
        for (int k = 0; k < 3; ++k) b.com[k] = b.trans[k]; // refreshed by the first substep
    }
    UPLOAD(dev.bodies, hb);
    std::vector<SimState> hs(1);
    hs[0].gravity[0] = params->gravity[0];
    hs[0].gravity[1] = params->gravity[1];
    hs[0].gravity[2] = (D == 3) ? params->gravity[2] : 0.0f;
    hs[0].dt = params->dt;
    hs[0].cell_width = cell_width;
    hs[0].num_bodies = (uint32_t)num_bodies;
    hs[0].slab_lo = -2147483647 - 1;
    hs[0].slab_hi = 2147483647;
    UPLOAD(dev.sim, hs);
    std::vector<Counters> hc(1);
    std::memset(hc.data(), 0, sizeof(Counters));
    hc[0].n_live = n;
    UPLOAD(dev.counters, hc);
#undef ALLOC
#undef UPLOAD
    if (num_bodies) launch_refresh_bodies(p->cfg(), dev); // world-space mass properties (rigid_impulses.wgsl:139-150)
    cudaError_t e = cudaStreamSynchronize(p->stream); // host vectors go out of scope
    if (e != cudaSuccess) {
        b200mpm_data_destroy(d);
        return fail(B200MPM_ERR_CUDA, cudaGetErrorString(e));
    }
    p->children.push_back(d);
    *out = d;
    return B200MPM_OK;
}

void b200mpm_data_destroy(b200mpm_data* d) {
    if (!d) return;
    if (d->pipe) {
        cudaSetDevice(d->pipe->device);
        cudaStreamSynchronize(d->pipe->stream);
        auto& ch = d->pipe->children;
        for (size_t i = 0; i < ch.size(); ++i)
            if (ch[i] == d) {
                ch.erase(ch.begin() + i);
                break;
            }
    } else {
        cudaSetDevice(d->device);
        cudaDeviceSynchronize();
    }
    for (auto& gp : d->graph_exec)
        for (auto& g : gp)
            if (g) cudaGraphExecDestroy(g);
    if (d->comm && nccl_api().ok) nccl_api().CommDestroy(d->comm);
    for (int k = 0; k < 2; ++k) {
        if (d->mig_send[k]) cudaFree(d->mig_send[k]);
        if (d->mig_recv[k]) cudaFree(d->mig_recv[k]);
        if (d->halo_send[k]) cudaFree(d->halo_send[k]);
        if (d->halo_recv[k]) cudaFree(d->halo_recv[k]);
    }
    if (d->imp_buf) cudaFree(d->imp_buf);
    for (int k = 0; k < 2; ++k)
        if (d->peer_arena[k]) cudaIpcCloseMemHandle(d->peer_arena[k]);
    if (d->arena) cudaFree(d->arena);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    for (int k = 0; k < 2; ++k) {
        if (d->pos_stage[k]) cudaFree(d->pos_stage[k]);
        if (d->pos_gathered[k]) cudaEventDestroy(d->pos_gathered[k]);
        if (d->pos_copied[k]) cudaEventDestroy(d->pos_copied[k]);
    }
    for (void* p : d->allocs) cudaFree(p);
    for (void* p : d->grid_allocs) cudaFree(p);
    if (d->staging) cudaFree(d->staging);
    if (d->pinned) cudaFreeHost(d->pinned);
    delete d;
}

size_t b200mpm_data_num_particles(const b200mpm_data* d) { return d ? d->n_live_host : 0; }
size_t b200mpm_data_num_bodies(const b200mpm_data* d) { return d ? d->num_bodies : 0; }

int b200mpm_data_set_rigid_particles(b200mpm_data* d, const float* vertices, const uint32_t* vertex_colliders,
                                     size_t num_vertices, const float* samples, const uint32_t* sample_ids,
                                     size_t num_samples) {
    if (!d) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null data");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    if (d->dev.num_rigid || d->dev.num_mesh_verts) return fail(B200MPM_ERR_INVALID_ARGUMENT, "rigid particles are already set");
    if (num_samples == 0) return B200MPM_OK;
    if (!vertices || !vertex_colliders || !samples || !sample_ids || num_vertices == 0)
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (num_samples >= (1ull << 31) || num_vertices >= (1ull << 31)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "too many points");
    const int prim = (p->dim == 3) ? 3 : 2;
    for (size_t i = 0; i < num_vertices; ++i)
        if (vertex_colliders[i] >= d->num_bodies) return fail(B200MPM_ERR_INVALID_ARGUMENT, "vertex refers to an unknown collider");
    for (size_t i = 0; i < num_samples; ++i) {
        if (sample_ids[4 * i + 3] >= d->num_bodies) return fail(B200MPM_ERR_INVALID_ARGUMENT, "sample refers to an unknown collider");
        for (int k = 0; k < prim; ++k)
            if (sample_ids[4 * i + k] >= num_vertices) return fail(B200MPM_ERR_INVALID_ARGUMENT, "sample refers to an unknown vertex");
    }
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaStreamSynchronize(p->stream));
    DeviceData& dev = d->dev;
    std::vector<float4> hv(num_vertices), hs(num_samples);
    std::vector<uint4> hi(num_samples);
    for (size_t i = 0; i < num_vertices; ++i) hv[i] = make_float4(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2], 0.f);
    for (size_t i = 0; i < num_samples; ++i) {
        hs[i] = make_float4(samples[3 * i], samples[3 * i + 1], samples[3 * i + 2], 0.f);
        hi[i] = make_uint4(sample_ids[4 * i], sample_ids[4 * i + 1], prim == 3 ? sample_ids[4 * i + 2] : 0u, sample_ids[4 * i + 3]);
    }
    int r = 0;
    if ((r = dev_alloc(d, &dev.mv_local, num_vertices))) return r;
    if ((r = dev_alloc(d, &dev.mv_world, num_vertices))) return r;
    if ((r = dev_alloc(d, &dev.mv_body, num_vertices))) return r;
    if ((r = dev_alloc(d, &dev.rp_local, num_samples))) return r;
    if ((r = dev_alloc(d, &dev.rp_world, num_samples))) return r;
    if ((r = dev_alloc(d, &dev.rp_ids, num_samples))) return r;
    if ((r = dev_alloc(d, &dev.rp_needs_block, num_samples))) return r;
    CU_TRY(cudaMemcpyAsync(dev.mv_local, hv.data(), num_vertices * sizeof(float4), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaMemcpyAsync(dev.mv_body, vertex_colliders, num_vertices * sizeof(uint32_t), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaMemcpyAsync(dev.rp_local, hs.data(), num_samples * sizeof(float4), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaMemcpyAsync(dev.rp_ids, hi.data(), num_samples * sizeof(uint4), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    dev.num_mesh_verts = (uint32_t)num_vertices;
    dev.num_rigid = (uint32_t)num_samples;
    for (auto& gp : d->graph_exec) // the captured substeps do not contain the rigid-particle kernels
        for (auto& g : gp)
            if (g) {
                cudaGraphExecDestroy(g);
                g = nullptr;
            }
    return B200MPM_OK;
}

int b200mpm_data_reserve_grid(b200mpm_data* d, uint32_t grid_capacity) {
    if (!d) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null data");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    if (grid_capacity == 0 || grid_capacity > (1u << 24))
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "grid_capacity must be in [1, 2^24]");
    uint32_t capacity = 1;
    while (capacity < grid_capacity) capacity <<= 1; // grid.rs:283
    if (capacity <= d->dev.capacity) return B200MPM_OK;
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaStreamSynchronize(p->stream));
    // The captured substep graphs hold the old pointers and the old capacity.
    for (auto& gp : d->graph_exec)
        for (auto& g : gp)
            if (g) {
                cudaGraphExecDestroy(g);
                g = nullptr;
            }
    // Allocate the new arrays first; the old grid is only given up once every allocation has succeeded.
    const DeviceData old_dev = d->dev;
    std::vector<void*> old_allocs;
    old_allocs.swap(d->grid_allocs);
    int r = alloc_grid(d, capacity);
    if (r) {
        for (void* q : d->grid_allocs) cudaFree(q);
        d->grid_allocs.swap(old_allocs);
        d->dev = old_dev;
        return r;
    }
    for (void* q : old_allocs) cudaFree(q);
    d->grid_dirty = true; // the new hash map is zero-filled, not cleared
    // Nothing of the old grid has to be cleared by the next k_begin_substep, and the (sticky) overflow flag is
    // about the capacity that was just replaced.
    const uint32_t zero = 0;
    CU_TRY(cudaMemcpyAsync(&d->dev.counters->prev_active_blocks, &zero, sizeof(zero), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaMemcpyAsync(&d->dev.counters->overflow, &zero, sizeof(zero), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

int b200mpm_data_set_auto_grow(b200mpm_data* d, float max_load) {
    if (!d) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null data");
    if (!(max_load >= 0.0f && max_load <= 1.0f)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "max_load must be in [0, 1]");
    d->auto_grow_load = max_load;
    return B200MPM_OK;
}

namespace {
// Auto-growth (off by default: the reference's capacity is fixed and its overflow silent). Looks at the block
// count the previous substeps left behind - one 4-byte readback, i.e. a stream synchronisation per step call.
int maybe_grow_grid(b200mpm_pipeline* p, b200mpm_data* d) {
    if (!(d->auto_grow_load > 0.0f)) return B200MPM_OK;
    uint32_t nb = 0;
    CU_TRY(cudaMemcpyAsync(&nb, &d->dev.counters->prev_active_blocks, sizeof(nb), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    uint64_t want = d->dev.capacity;
    while ((double)nb > (double)d->auto_grow_load * (double)want && want < (1ull << 24)) want <<= 1;
    if (want > d->dev.capacity) return b200mpm_data_reserve_grid(d, (uint32_t)want);
    return B200MPM_OK;
}
} // namespace

int b200mpm_step(b200mpm_pipeline* p, b200mpm_data* d, uint32_t num_substeps) {
    if (!p || !d || d->pipe != p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "pipeline/data mismatch");
    CU_TRY(cudaSetDevice(p->device));
    {
        int r = maybe_grow_grid(p, d);
        if (r) return r;
    }
    for (uint32_t s = 0; s < num_substeps; ++s) run_substep(p, d);
    launch_integrate_bodies(p->cfg(), d->dev); // flushes the integration the last graph replay deferred (else a no-op)
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_sort_only(b200mpm_pipeline* p, b200mpm_data* d) {
    if (!p || !d || d->pipe != p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "pipeline/data mismatch");
    CU_TRY(cudaSetDevice(p->device));
    LaunchCfg c = p->cfg();
    ensure_clean_grid(p, d);
    launch_transform_rigid(c, d->dev);
    run_sort(p, d);
    d->grid_dirty = true;
    d->sorted_indirect = true;
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_sync(b200mpm_pipeline* p) {
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null pipeline");
    CU_TRY(cudaSetDevice(p->device));
    CU_TRY(cudaStreamSynchronize(p->stream));
    for (b200mpm_data* d : p->children) // pending asynchronous readbacks
        if (d->copy_stream) CU_TRY(cudaStreamSynchronize(d->copy_stream));
    return B200MPM_OK;
}

int b200mpm_set_timestamps(b200mpm_pipeline* p, int enabled) {
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null pipeline");
    if (!enabled) fold_events(p);
    p->timestamps = enabled != 0;
    return B200MPM_OK;
}

int b200mpm_debug_timeline(b200mpm_data* d, uint64_t ns[2 * B200MPM_NUM_KERNELS]) {
    if (!d) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    const bool want = ns != nullptr;
    uint64_t init[2 * B200MPM_NUM_KERNELS];
    for (int k = 0; k < B200MPM_NUM_KERNELS; ++k) init[2 * k] = ~0ull, init[2 * k + 1] = 0ull;
    if (want != (d->dev.timeline_on != 0)) { // the switch is a kernel argument: the captured substeps are stale
        CU_TRY(cudaStreamSynchronize(p->stream));
        for (auto& gp : d->graph_exec)
            for (auto& g : gp)
                if (g) {
                    cudaGraphExecDestroy(g);
                    g = nullptr;
                }
        d->dev.timeline_on = want ? 1 : 0;
        if (want) std::memcpy(ns, init, sizeof(init)); // nothing recorded yet
    } else if (want) {
        CU_TRY(cudaMemcpyAsync(ns, d->dev.timeline, sizeof(uint64_t) * 2 * B200MPM_NUM_KERNELS, cudaMemcpyDeviceToHost, p->stream));
        CU_TRY(cudaStreamSynchronize(p->stream));
    }
    if (!want) return B200MPM_OK;
    CU_TRY(cudaMemcpyAsync(d->dev.timeline, init, sizeof(init), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

int b200mpm_get_kernel_timings(b200mpm_pipeline* p, double ms[B200MPM_NUM_KERNELS]) {
    if (!p || !ms) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    fold_events(p);
    for (int i = 0; i < B200MPM_NUM_KERNELS; ++i) {
        ms[i] = p->kernel_ms[i];
        p->kernel_ms[i] = 0.0;
    }
    return B200MPM_OK;
}

int b200mpm_get_timings(b200mpm_pipeline* p, double ms[B200MPM_NUM_PASSES]) {
    if (!p || !ms) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    fold_events(p);
    for (int i = 0; i < B200MPM_NUM_PASSES; ++i) {
        ms[i] = p->pass_ms[i];
        p->pass_ms[i] = 0.0;
    }
    return B200MPM_OK;
}

int b200mpm_write_sim_params(b200mpm_data* d, const b200mpm_sim_params* params) {
    if (!d || !params || !d->pipe) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument / destroyed pipeline");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    int r = ensure_pinned(d, 4096);
    if (r) return r;
    CU_TRY(cudaStreamSynchronize(p->stream)); // the pinned mirror may still be in flight
    float* h = (float*)d->pinned;
    h[0] = params->gravity[0];
    h[1] = params->gravity[1];
    h[2] = (p->dim == 3) ? params->gravity[2] : 0.0f;
    h[3] = params->dt;
    CU_TRY(cudaMemcpyAsync(d->dev.sim, h, 4 * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    return B200MPM_OK;
}

static_assert(B200MPM_MAX_BODIES * sizeof(b200mpm_pose) <= 2048 && B200MPM_MAX_BODIES * sizeof(b200mpm_velocity) <= 2048,
              "body uploads use 2 KB halves of the pinned / staging areas");

int b200mpm_write_body_poses(b200mpm_data* d, const b200mpm_pose* poses, size_t n) {
    if (!d || (!poses && n)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (n > d->num_bodies) return fail(B200MPM_ERR_INVALID_ARGUMENT, "more poses than bodies");
    if (n == 0) return B200MPM_OK;
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    int r = ensure_pinned(d, 4096);
    if (r) return r;
    r = ensure_staging(d, 4096);
    if (r) return r;
    // Poses and velocities use disjoint halves of the 4 KB pinned / staging areas, and the synchronisation in
    // front of the host copy covers the previous use of this half: no trailing synchronisation, the upload
    // overlaps with the host preparing the next call.
    CU_TRY(cudaStreamSynchronize(p->stream));
    std::memcpy(d->pinned, poses, n * sizeof(b200mpm_pose));
    CU_TRY(cudaMemcpyAsync(d->staging, d->pinned, n * sizeof(b200mpm_pose), cudaMemcpyHostToDevice, p->stream));
    launch_write_poses(p->cfg(), d->dev, (const b200mpm_pose*)d->staging, (uint32_t)n);
    return B200MPM_OK;
}

int b200mpm_write_body_vels(b200mpm_data* d, const b200mpm_velocity* vels, size_t n) {
    if (!d || (!vels && n)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (n > d->num_bodies) return fail(B200MPM_ERR_INVALID_ARGUMENT, "more velocities than bodies");
    if (n == 0) return B200MPM_OK;
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    int r = ensure_pinned(d, 4096);
    if (r) return r;
    r = ensure_staging(d, 4096);
    if (r) return r;
    char* pinned_half = (char*)d->pinned + 2048;
    char* staging_half = (char*)d->staging + 2048;
    CU_TRY(cudaStreamSynchronize(p->stream));
    std::memcpy(pinned_half, vels, n * sizeof(b200mpm_velocity));
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
            if (vels[i].linear[k] != 0.0f || vels[i].angular[k] != 0.0f) {
                if (!d->bodies_react) // the captured substeps have neither the impulse pass nor its all-reduce yet
                    for (auto& gp : d->graph_exec)
                        for (auto& g : gp)
                            if (g) {
                                cudaGraphExecDestroy(g);
                                g = nullptr;
                            }
                d->bodies_react = true;
                d->dev.bodies_react = 1;
            }
    CU_TRY(cudaMemcpyAsync(staging_half, pinned_half, n * sizeof(b200mpm_velocity), cudaMemcpyHostToDevice, p->stream));
    launch_write_vels(p->cfg(), d->dev, (const b200mpm_velocity*)staging_half, (uint32_t)n);
    return B200MPM_OK;
}

static int read_body_state(b200mpm_data* d, b200mpm_pose* poses, b200mpm_velocity* vels, size_t n) {
    if (!d) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (n > d->num_bodies) n = d->num_bodies;
    if (n == 0) return B200MPM_OK;
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    int r = ensure_pinned(d, 4096);
    if (r) return r;
    r = ensure_staging(d, 4096);
    if (r) return r;
    b200mpm_pose* dp = (b200mpm_pose*)d->staging;
    b200mpm_velocity* dv = (b200mpm_velocity*)((char*)d->staging + 2048);
    launch_read_poses(p->cfg(), d->dev, poses ? dp : nullptr, vels ? dv : nullptr, (uint32_t)n);
    CU_TRY(cudaMemcpyAsync(d->pinned, d->staging, 4096, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    if (poses) std::memcpy(poses, d->pinned, n * sizeof(b200mpm_pose));
    if (vels) std::memcpy(vels, (char*)d->pinned + 2048, n * sizeof(b200mpm_velocity));
    return B200MPM_OK;
}
int b200mpm_read_body_poses(b200mpm_data* d, b200mpm_pose* poses, size_t n) {
    if (!poses && n) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    return read_body_state(d, poses, nullptr, n);
}
int b200mpm_read_body_vels(b200mpm_data* d, b200mpm_velocity* vels, size_t n) {
    if (!vels && n) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    return read_body_state(d, nullptr, vels, n);
}

int b200mpm_read_positions(b200mpm_data* d, float* out) {
    if (!d || (!out && d->dev.n)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (d->dev.n == 0) return B200MPM_OK;
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    size_t bytes = (size_t)d->dev.n * sizeof(float4);
    int r = ensure_staging(d, bytes);
    if (r) return r;
    if (d->sharded) return fail(B200MPM_ERR_INVALID_ARGUMENT, "sharded data: use b200mpm_read_particles_unordered");
    if (!d->ordered_readback_ok) return fail(B200MPM_ERR_INVALID_ARGUMENT, "created with custom particle ids or spare capacity: use the unordered read-backs");
    launch_gather_positions(p->cfg(), d->dev, d->cur, (float4*)d->staging);
    CU_TRY(cudaMemcpyAsync(out, d->staging, bytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

namespace {
// Two device staging slots + a copy stream: the gather of readback k+1 may run while copy k is still in flight.
int ensure_async_readback(b200mpm_data* d) {
    if (d->copy_stream) return B200MPM_OK;
    const size_t bytes = (size_t)d->dev.n * sizeof(float4);
    CU_TRY(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
        CU_TRY(cudaMalloc(&d->pos_stage[k], bytes));
        CU_TRY(cudaEventCreateWithFlags(&d->pos_gathered[k], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&d->pos_copied[k], cudaEventDisableTiming));
        CU_TRY(cudaEventRecord(d->pos_copied[k], d->copy_stream));
    }
    return B200MPM_OK;
}
int enqueue_async_positions(b200mpm_pipeline* p, b200mpm_data* d, float* out, size_t count, int unordered) {
    const int slot = d->pos_slot ^= 1;
    // the slot's previous copy must have left the device before the gather overwrites it
    CU_TRY(cudaStreamWaitEvent(p->stream, d->pos_copied[slot], 0));
    launch_gather_positions(p->cfg(), d->dev, d->cur, d->pos_stage[slot], unordered);
    CU_TRY(cudaEventRecord(d->pos_gathered[slot], p->stream));
    CU_TRY(cudaStreamWaitEvent(d->copy_stream, d->pos_gathered[slot], 0));
    CU_TRY(cudaMemcpyAsync(out, d->pos_stage[slot], count * sizeof(float4), cudaMemcpyDeviceToHost, d->copy_stream));
    CU_TRY(cudaEventRecord(d->pos_copied[slot], d->copy_stream));
    return B200MPM_OK;
}
} // namespace

int b200mpm_read_positions_async(b200mpm_data* d, float* out) {
    if (!d || (!out && d->dev.n)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (d->dev.n == 0) return B200MPM_OK;
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    if (d->sharded) return fail(B200MPM_ERR_INVALID_ARGUMENT, "sharded data: use b200mpm_read_positions_unordered_async");
    if (!d->ordered_readback_ok) return fail(B200MPM_ERR_INVALID_ARGUMENT, "created with custom particle ids or spare capacity: use the unordered read-backs");
    CU_TRY(cudaSetDevice(p->device));
    int r = ensure_async_readback(d);
    if (r) return r;
    return enqueue_async_positions(p, d, out, d->dev.n, 0);
}

int b200mpm_read_positions_unordered_async(b200mpm_data* d, float* out, size_t capacity, size_t* count) {
    if (!d || !count) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    // No host synchronisation: the copy has a FIXED size - every slot of the particle capacity, the spare ones marked
    // "no particle" (id NONE) by the gather kernel - so it does not have to wait for the enqueued substeps to learn the
    // live count (which used to cost one stream synchronisation per frame and rank, i.e. no overlap at all).
    const size_t n = d->dev.n;
    *count = n;
    if (n == 0) return B200MPM_OK;
    if (!out || capacity < n) return fail(B200MPM_ERR_INVALID_ARGUMENT, "output too small: it must hold particle_capacity entries");
    int r = ensure_async_readback(d);
    if (r) return r;
    return enqueue_async_positions(p, d, out, n, 1);
}

int b200mpm_prep_vertex_buffer(b200mpm_pipeline* p, b200mpm_data* d, b200mpm_instance* dev_instances, uint32_t mode) {
    if (!p || !d || d->pipe != p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "pipeline/data mismatch");
    if (!dev_instances && d->dev.n) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null instance buffer");
    if (mode > B200MPM_RENDER_CDF_SIGNS) return fail(B200MPM_ERR_INVALID_ARGUMENT, "unknown render mode");
    if (d->sharded || !d->ordered_readback_ok)
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "sharded data / custom particle ids / spare capacity: the instance buffer is indexed by the default particle ids");
    CU_TRY(cudaSetDevice(p->device));
    cudaPointerAttributes attr{};
    if (d->dev.n && (cudaPointerGetAttributes(&attr, dev_instances) != cudaSuccess ||
                     (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged))) {
        cudaGetLastError();
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "the instance buffer must be device memory");
    }
    launch_prep_vertex_buffer(p->cfg(), d->dev, d->cur, dev_instances, mode);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_read_positions_unordered(b200mpm_data* d, float* out, size_t capacity, size_t* count) {
    if (!d || !count) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    int r = ensure_staging(d, (size_t)d->dev.n * sizeof(float4) + 256);
    if (r) return r;
    launch_gather_positions(p->cfg(), d->dev, d->cur, (float4*)d->staging, 1);
    Counters c;
    CU_TRY(cudaMemcpyAsync(&c, d->dev.counters, sizeof(c), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    *count = c.n_live;
    if (c.n_live == 0) return B200MPM_OK;
    if (!out || capacity < c.n_live) return fail(B200MPM_ERR_INVALID_ARGUMENT, "output too small");
    CU_TRY(cudaMemcpyAsync(out, d->staging, (size_t)c.n_live * sizeof(float4), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

int b200mpm_read_particles(b200mpm_data* d, b200mpm_particle* out) {
    if (!d || (!out && d->dev.n)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (d->dev.n == 0) return B200MPM_OK;
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    size_t bytes = (size_t)d->dev.n * sizeof(b200mpm_particle);
    int r = ensure_staging(d, bytes);
    if (r) return r;
    if (d->sharded) return fail(B200MPM_ERR_INVALID_ARGUMENT, "sharded data: use b200mpm_read_particles_unordered");
    if (!d->ordered_readback_ok) return fail(B200MPM_ERR_INVALID_ARGUMENT, "created with custom particle ids or spare capacity: use the unordered read-backs");
    launch_gather_particles(p->cfg(), d->dev, d->cur, (b200mpm_particle*)d->staging, nullptr);
    CU_TRY(cudaMemcpyAsync(out, d->staging, bytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

int b200mpm_data_status(b200mpm_data* d, uint32_t* num_active_blocks) {
    if (!d) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    Counters c;
    CU_TRY(cudaMemcpyAsync(&c, d->dev.counters, sizeof(c), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    if (num_active_blocks) *num_active_blocks = std::min(c.prev_active_blocks, d->dev.capacity);
    // (bit 2: live particles were left without a block; codes 1..3 in the low bits)
    if ((c.overflow & 4u) && d->sharded)
        return fail(B200MPM_ERR_GRID_OVERFLOW, "grid block capacity exceeded: blocks were dropped and their particles lost (sharded run)");
    if ((c.overflow & 3u) == 1 || c.overflow == 4u) return fail(B200MPM_ERR_GRID_OVERFLOW, "grid block capacity exceeded: blocks were dropped");
    if ((c.overflow & 3u) == 2) return fail(B200MPM_ERR_GRID_OVERFLOW, "a shard exchange buffer was too small");
    if ((c.overflow & 3u) == 3) return fail(B200MPM_ERR_GRID_OVERFLOW, "particle_capacity exceeded by immigrants: particles were lost");
    return B200MPM_OK;
}

int b200mpm_read_grid(b200mpm_data* d, b200mpm_block_info* blocks, b200mpm_node* nodes, size_t capacity,
                      size_t* num_blocks) {
    if (!d || !num_blocks) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    uint32_t nb = 0;
    int st = b200mpm_data_status(d, &nb);
    if (st != B200MPM_OK && st != B200MPM_ERR_GRID_OVERFLOW) return st;
    size_t take = std::min<size_t>(nb, capacity);
    *num_blocks = take;
    if (take == 0) return B200MPM_OK;
    if (!blocks || !nodes) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null output");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    size_t bbytes = take * sizeof(b200mpm_block_info);
    size_t boff = (bbytes + 255) & ~(size_t)255;
    size_t nbytes = take * CELLS_PER_BLOCK * sizeof(b200mpm_node);
    int r = ensure_staging(d, boff + nbytes);
    if (r) return r;
    b200mpm_block_info* db = (b200mpm_block_info*)d->staging;
    b200mpm_node* dn = (b200mpm_node*)((char*)d->staging + boff);
    launch_gather_grid(p->cfg(), d->dev, db, dn, (uint32_t)take);
    CU_TRY(cudaMemcpyAsync(blocks, db, bbytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaMemcpyAsync(nodes, dn, nbytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

int b200mpm_read_sorted_ids(b200mpm_data* d, uint32_t* out) {
    if (!d || (!out && d->dev.n)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (d->dev.n == 0) return B200MPM_OK;
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    if (!d->ordered_readback_ok || d->sharded)
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "created with custom particle ids or spare capacity: the sorted-id read-back needs the default layout");
    size_t bytes = (size_t)d->dev.n * sizeof(uint32_t);
    int r = ensure_staging(d, bytes);
    if (r) return r;
    launch_gather_sorted_ids(p->cfg(), d->dev, d->cur, d->sorted_indirect ? 1 : 0, (uint32_t*)d->staging);
    CU_TRY(cudaMemcpyAsync(out, d->staging, bytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

int b200mpm_prefix_sum_u32(b200mpm_pipeline* p, uint32_t* data, size_t len) {
    if (!p || (!data && len)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (len == 0) return B200MPM_OK;
    if (len >= (1ull << 31)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "vector too long");
    CU_TRY(cudaSetDevice(p->device));
    uint32_t* dd = nullptr;
    uint64_t* state = nullptr;
    uint32_t* ticket = nullptr;
    uint32_t tiles = scan_num_tiles(len);
    CU_TRY(cudaMalloc(&dd, len * sizeof(uint32_t)));
    cudaError_t e1 = cudaMalloc(&state, (tiles + 2) * sizeof(uint64_t));
    cudaError_t e2 = cudaMalloc(&ticket, sizeof(uint32_t));
    int rc = B200MPM_OK;
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        rc = fail(B200MPM_ERR_OUT_OF_MEMORY, "cudaMalloc failed");
    } else {
        cudaMemcpyAsync(dd, data, len * sizeof(uint32_t), cudaMemcpyHostToDevice, p->stream);
        launch_exclusive_scan_u32(p->cfg(), dd, (uint32_t)len, state, ticket);
        cudaMemcpyAsync(data, dd, len * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream);
        cudaError_t e = cudaStreamSynchronize(p->stream);
        if (e != cudaSuccess) rc = fail(B200MPM_ERR_CUDA, cudaGetErrorString(e));
    }
    cudaFree(dd);
    if (state) cudaFree(state);
    if (ticket) cudaFree(ticket);
    return rc;
}

int b200mpm_slab_configure(b200mpm_data* d, int32_t x_lo, int32_t x_hi) {
    if (!d) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (x_lo >= x_hi) return fail(B200MPM_ERR_INVALID_ARGUMENT, "empty slab");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    int r = ensure_pinned(d, 4096);
    if (r) return r;
    CU_TRY(cudaStreamSynchronize(p->stream));
    int* h = (int*)d->pinned;
    h[0] = x_lo;
    h[1] = x_hi;
    CU_TRY(cudaMemcpyAsync(&d->dev.sim->slab_lo, h, 2 * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    d->sharded = true;
    return B200MPM_OK;
}

int b200mpm_data_num_live(b200mpm_data* d, size_t* num_live) {
    if (!d || !num_live) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    CU_TRY(cudaSetDevice(p->device));
    Counters c;
    CU_TRY(cudaMemcpyAsync(&c, d->dev.counters, sizeof(c), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    d->n_live_host = c.n_live;
    *num_live = c.n_live;
    return B200MPM_OK;
}

int b200mpm_shard_emigrate(b200mpm_pipeline* p, b200mpm_data* d, void* dev_left, void* dev_right, uint32_t cap_records) {
    if (!p || !d || d->pipe != p || !dev_left || !dev_right) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    launch_emigrate(p->cfg(), d->dev, d->cur, dev_left, dev_right, cap_records);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_shard_immigrate(b200mpm_pipeline* p, b200mpm_data* d, const void* dev_buffer, uint32_t cap_records) {
    if (!p || !d || d->pipe != p || !dev_buffer) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    launch_immigrate(p->cfg(), d->dev, d->cur, dev_buffer, cap_records);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_shard_step_begin(b200mpm_pipeline* p, b200mpm_data* d) {
    if (!p || !d || d->pipe != p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "pipeline/data mismatch");
    CU_TRY(cudaSetDevice(p->device));
    run_phase(p, d, PHASE_BEGIN);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_shard_step_end(b200mpm_pipeline* p, b200mpm_data* d) {
    if (!p || !d || d->pipe != p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "pipeline/data mismatch");
    CU_TRY(cudaSetDevice(p->device));
    run_phase(p, d, PHASE_END);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_shard_halo_pack(b200mpm_pipeline* p, b200mpm_data* d, void* dev_left, void* dev_right, uint32_t cap_blocks) {
    if (!p || !d || d->pipe != p || !dev_left || !dev_right) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    launch_halo_pack(p->cfg(), d->dev, dev_left, dev_right, cap_blocks);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_shard_halo_add(b200mpm_pipeline* p, b200mpm_data* d, const void* dev_buffer, uint32_t cap_blocks) {
    if (!p || !d || d->pipe != p || !dev_buffer) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    launch_halo_add(p->cfg(), d->dev, dev_buffer, nullptr, cap_blocks);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_shard_impulses(b200mpm_pipeline* p, b200mpm_data* d, int32_t* dev_buf, int write) {
    if (!p || !d || d->pipe != p || !dev_buf) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    CU_TRY(cudaSetDevice(p->device));
    launch_impulses_io(p->cfg(), d->dev, dev_buf, write);
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_nccl_unique_id(void* out, size_t bytes) {
    if (!out || bytes < sizeof(ncclUniqueId)) return fail(B200MPM_ERR_INVALID_ARGUMENT, "need a 128-byte buffer");
    const NcclApi& nc = nccl_api();
    if (!nc.ok) return fail(B200MPM_ERR_COMM, nc.error);
    ncclUniqueId id;
    ncclResult_t r = nc.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(B200MPM_ERR_COMM, nc.GetErrorString(r));
    std::memcpy(out, &id, sizeof(id));
    return B200MPM_OK;
}

int b200mpm_shard_comm_init(b200mpm_pipeline* p, b200mpm_data* d, int rank, int world, const void* unique_id,
                            uint32_t migration_cap, uint32_t halo_cap) {
    if (!p || !d || d->pipe != p || !unique_id) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (world < 1 || rank < 0 || rank >= world || !migration_cap || !halo_cap)
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "bad rank / world / capacities");
    if (d->comm) return fail(B200MPM_ERR_INVALID_ARGUMENT, "communicator already initialised");
    const NcclApi& nc = nccl_api();
    if (!nc.ok) return fail(B200MPM_ERR_COMM, nc.error);
    CU_TRY(cudaSetDevice(p->device));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    ncclResult_t r = nc.CommInitRank(&d->comm, world, id, rank);
    if (r != ncclSuccess) return fail(B200MPM_ERR_COMM, nc.GetErrorString(r));
    d->rank = rank;
    d->world = world;
    d->mig_cap = migration_cap;
    d->halo_cap = halo_cap;
    const size_t mig_bytes = B200MPM_SHARD_HEADER_BYTES + (size_t)migration_cap * B200MPM_PARTICLE_RECORD_BYTES;
    const size_t halo_bytes = B200MPM_SHARD_HEADER_BYTES + (size_t)halo_cap * B200MPM_HALO_BLOCK_BYTES;
    for (int k = 0; k < 2; ++k) {
        CU_TRY(cudaMalloc(&d->mig_send[k], mig_bytes));
        CU_TRY(cudaMalloc(&d->mig_recv[k], mig_bytes));
        CU_TRY(cudaMalloc(&d->halo_send[k], halo_bytes));
        CU_TRY(cudaMalloc(&d->halo_recv[k], halo_bytes));
        CU_TRY(cudaMemsetAsync(d->mig_send[k], 0, mig_bytes, p->stream));
        CU_TRY(cudaMemsetAsync(d->mig_recv[k], 0, mig_bytes, p->stream));
        CU_TRY(cudaMemsetAsync(d->halo_send[k], 0, halo_bytes, p->stream));
        CU_TRY(cudaMemsetAsync(d->halo_recv[k], 0, halo_bytes, p->stream));
    }
    CU_TRY(cudaMalloc(&d->imp_buf, B200MPM_MAX_BODIES * 6 * sizeof(int)));
    CU_TRY(cudaStreamSynchronize(p->stream));
    d->sharded = true;
    return B200MPM_OK;
}

int b200mpm_shard_p2p_export(b200mpm_pipeline* p, b200mpm_data* d, void* handle_out, size_t bytes) {
    if (!p || !d || d->pipe != p || !handle_out || bytes < sizeof(cudaIpcMemHandle_t))
        return fail(B200MPM_ERR_INVALID_ARGUMENT, "need a 64-byte handle buffer");
    if (!d->comm) return fail(B200MPM_ERR_INVALID_ARGUMENT, "b200mpm_shard_comm_init has not been called");
    if (d->arena) return fail(B200MPM_ERR_INVALID_ARGUMENT, "arena already exported");
    CU_TRY(cudaSetDevice(p->device));
    d->arena_mig_bytes = (B200MPM_SHARD_HEADER_BYTES + (size_t)d->mig_cap * B200MPM_PARTICLE_RECORD_BYTES + 255) & ~(size_t)255;
    d->arena_halo_bytes = (B200MPM_SHARD_HEADER_BYTES + (size_t)d->halo_cap * B200MPM_HALO_BLOCK_BYTES + 255) & ~(size_t)255;
    const size_t total = 256 + 4 * d->arena_mig_bytes + 4 * d->arena_halo_bytes;
    CU_TRY(cudaMalloc((void**)&d->arena, total));
    CU_TRY(cudaMemset(d->arena, 0, total));
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, d->arena));
    std::memcpy(handle_out, &h, sizeof(h));
    return B200MPM_OK;
}

int b200mpm_shard_p2p_connect(b200mpm_pipeline* p, b200mpm_data* d, const void* handles, size_t num_handles) {
    if (!p || !d || d->pipe != p || !handles) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    if (!d->arena || (int)num_handles != d->world) return fail(B200MPM_ERR_INVALID_ARGUMENT, "export first; one handle per rank");
    CU_TRY(cudaSetDevice(p->device));
    const cudaIpcMemHandle_t* hs = (const cudaIpcMemHandle_t*)handles;
    const int nbr[2] = {d->rank - 1, d->rank + 1};
    for (int k = 0; k < 2; ++k) {
        if (nbr[k] < 0 || nbr[k] >= d->world) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hs + nbr[k], sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < k; ++j)
                if (d->peer_arena[j]) {
                    cudaIpcCloseMemHandle(d->peer_arena[j]);
                    d->peer_arena[j] = nullptr;
                }
            return fail(B200MPM_ERR_COMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
        }
        d->peer_arena[k] = (char*)ptr;
    }
    for (auto& gp : d->graph_exec)
        for (auto& g : gp)
            if (g) { // (k_g2p's argument block changes below: every captured substep is stale)
                cudaGraphExecDestroy(g);
                g = nullptr;
            }
    // From here on k_g2p lists the particles that leave the slab, and the migration packs that list instead of
    // scanning the slab (shard.cu); whoever is outside already is listed once, now.
    if (!d->dev.emig_list) {
        d->dev.emig_cap = 2 * d->mig_cap;
        int r = dev_alloc(d, &d->dev.emig_list, d->dev.emig_cap);
        if (r) return r;
        launch_list_emigrants(p->cfg(), d->dev, d->cur);
        CU_TRY(cudaStreamSynchronize(p->stream));
    }
    d->p2p_ready = true;
    return B200MPM_OK;
}

int b200mpm_shard_step(b200mpm_pipeline* p, b200mpm_data* d, uint32_t num_substeps) {
    if (!p || !d || d->pipe != p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "pipeline/data mismatch");
    if (!d->comm) return fail(B200MPM_ERR_INVALID_ARGUMENT, "b200mpm_shard_comm_init has not been called");
    CU_TRY(cudaSetDevice(p->device));
    {
        int r = maybe_grow_grid(p, d); // per rank: the capacity is a local property of each slab
        if (r) return r;
    }
    for (uint32_t s = 0; s < num_substeps; ++s) run_phase(p, d, PHASE_SHARDED);
    launch_integrate_bodies(p->cfg(), d->dev); // flushes the integration the last graph replay deferred (else a no-op)
    CU_TRY(cudaGetLastError());
    return B200MPM_OK;
}

int b200mpm_read_particles_unordered(b200mpm_data* d, b200mpm_particle* out, uint32_t* ids, size_t capacity,
                                     size_t* count) {
    if (!d || !count) return fail(B200MPM_ERR_INVALID_ARGUMENT, "null argument");
    size_t live = 0;
    int r = b200mpm_data_num_live(d, &live);
    if (r) return r;
    *count = live;
    if (live == 0) return B200MPM_OK;
    if (!out || !ids || capacity < live) return fail(B200MPM_ERR_INVALID_ARGUMENT, "output too small");
    b200mpm_pipeline* p = d->pipe;
    if (!p) return fail(B200MPM_ERR_INVALID_ARGUMENT, "the pipeline of this data object was destroyed");
    size_t pbytes = live * sizeof(b200mpm_particle);
    size_t poff = ((size_t)d->dev.n * sizeof(b200mpm_particle) + 255) & ~(size_t)255;
    r = ensure_staging(d, poff + (size_t)d->dev.n * sizeof(uint32_t));
    if (r) return r;
    uint32_t* dids = (uint32_t*)((char*)d->staging + poff);
    launch_gather_particles(p->cfg(), d->dev, d->cur, (b200mpm_particle*)d->staging, dids);
    CU_TRY(cudaMemcpyAsync(out, d->staging, pbytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaMemcpyAsync(ids, dids, live * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(cudaStreamSynchronize(p->stream));
    return B200MPM_OK;
}

} // extern "C"
