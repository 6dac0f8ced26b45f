// Multi-GPU slab sharding (no counterpart in the reference, which is single-device).
//
// The domain is cut into slabs of grid-block columns along x; rank r owns the particles whose block
// x-index lies in [slab_lo, slab_hi). Per substep:
//   1. migration   — particles whose block left the slab are packed for the -x / +x neighbour and flagged
//                    dead; immigrants are appended to the live range;
//   2. P2G         — on the rank's own particles only; the block column `slab_hi` (written by the last owned
//                    column, owned by the +x neighbour) and the column `slab_lo` hold PARTIAL node sums;
//   3. node halo   — both neighbours exchange their partial sums of the shared column (keyed by the block's
//                    virtual id) and add them: a + b on one side, b + a on the other, bitwise the same;
//   4. impulses    — the 16 x 6 fixed-point body impulses are summed over ranks (exact, integers);
//   5. G2P + update, body integration (replicated, deterministic).
// The kernels here only pack / unpack device buffers; the transport (NCCL send/recv, all-reduce) belongs to the
// host (wgsparkl_b200/sharded.py with torch.distributed; a Rust host would call NCCL directly).
#include "launch.h"

namespace b2 {

// One migrating particle: 8 x 16 bytes.
struct __align__(16) ParticleRecord {
    float4 pos4, vel4, Fa, Fb, Ca, Cb;
    float Fc, Cc;
    uint32_t cdf_aff, pad;
    float4 plastic;
};
static_assert(sizeof(ParticleRecord) == 128, "ParticleRecord must be 128 bytes");

// One halo block: virtual id + 64 nodes (momentum xyz + mass).
struct __align__(16) HaloBlock {
    int4 vid;
    float4 node[CELLS_PER_BLOCK];
};
static_assert(sizeof(HaloBlock) == 16 + 1024, "HaloBlock must be 1040 bytes");

// Every exchange buffer starts with a 16-byte header holding the record count, so that neither side needs a
// host round trip: fixed-size buffers travel, the receiver reads the count on the device.
struct __align__(16) ShardHeader {
    uint32_t count, pad0, pad1, pad2;
};
template <class T>
__device__ __forceinline__ T* shard_records(void* buf) { return (T*)((char*)buf + sizeof(ShardHeader)); }
template <class T>
__device__ __forceinline__ const T* shard_records(const void* buf) { return (const T*)((const char*)buf + sizeof(ShardHeader)); }

// Packs particle i for the neighbour it left the slab towards (if any) and flags it dead.
template <int D>
__device__ __forceinline__ void emigrate_particle(const DeviceData& d, int cur, uint32_t i, ParticleRecord* left,
                                                  ParticleRecord* right, uint32_t cap) {
    float4 p = d.pos4[cur][i];
    if (__float_as_uint(p.w) & FLAG_DEAD) return;
    const int bx = assoc_cell(p.x, d.sim->cell_width, 1.0f / d.sim->cell_width) >> Dim<D>::LOG_BLOCK;
    const int dir = (bx < d.sim->slab_lo) ? 0 : (bx >= d.sim->slab_hi) ? 1 : -1;
    if (dir < 0) return;
    const uint32_t slot = atomicAdd(&d.counters->send_count[dir], 1u);
    if (slot >= cap) { // buffer too small: the particle stays (and is retried next substep); reported via status
        d.counters->overflow = 2u;
        return;
    }
    ParticleRecord r;
    r.pos4 = p;
    r.vel4 = d.vel4[cur][i];
    r.Fa = d.Fa[cur][i];
    r.Ca = d.Ca[cur][i];
    r.Fb = make_float4(0.f, 0.f, 0.f, 0.f);
    r.Cb = r.Fb;
    r.Fc = r.Cc = 0.0f;
    if (D == 3) {
        r.Fb = d.Fb[cur][i];
        r.Cb = d.Cb[cur][i];
        r.Fc = d.Fc[cur][i];
        r.Cc = d.Cc[cur][i];
    }
    r.cdf_aff = d.has_bodies ? d.cdf_aff[cur][i] : 0u;
    r.pad = 0u;
    r.plastic = d.has_plastic ? d.plastic[cur][i] : make_float4(1.0f, 1.0f, 0.0f, 0.0f);
    (dir == 0 ? left : right)[slot] = r;
    d.pos4[cur][i].w = __uint_as_float(__float_as_uint(p.w) | FLAG_DEAD);
}

template <int D>
__global__ void __launch_bounds__(256) k_emigrate(DeviceData d, int cur, void* left_buf, void* right_buf, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.counters->n_live) return;
    emigrate_particle<D>(d, cur, i, shard_records<ParticleRecord>(left_buf), shard_records<ParticleRecord>(right_buf), cap);
}

// One-time scan (when the peer-to-peer path is switched on): whoever is outside the slab already goes on the list
// that k_g2p keeps from then on.
template <int D>
__global__ void __launch_bounds__(256) k_list_emigrants(DeviceData d, int cur) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.counters->n_live) return;
    const float4 p = d.pos4[cur][i];
    if (__float_as_uint(p.w) & FLAG_DEAD) return;
    const int bx = assoc_cell(p.x, d.sim->cell_width, 1.0f / d.sim->cell_width) >> Dim<D>::LOG_BLOCK;
    if (bx < d.sim->slab_lo || bx >= d.sim->slab_hi) {
        const uint32_t slot = atomicAdd(&d.counters->emig_count, 1u);
        if (slot < d.emig_cap) d.emig_list[slot] = i;
    }
}

template <int D>
__global__ void __launch_bounds__(256) k_immigrate(DeviceData d, int cur, const void* buf, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = min(((const ShardHeader*)buf)->count, cap);
    if (i >= count) return;
    const uint32_t dst = d.counters->n_live + i; // n_live is bumped by k_immigrate_commit afterwards
    if (dst >= d.n) {
        d.counters->overflow = 3u; // particle_capacity exceeded: the immigrant is lost
        return;
    }
    const ParticleRecord r = shard_records<ParticleRecord>(buf)[i];
    d.pos4[cur][dst] = r.pos4;
    d.vel4[cur][dst] = r.vel4;
    d.Fa[cur][dst] = r.Fa;
    d.Ca[cur][dst] = r.Ca;
    if (D == 3) {
        d.Fb[cur][dst] = r.Fb;
        d.Cb[cur][dst] = r.Cb;
        d.Fc[cur][dst] = r.Fc;
        d.Cc[cur][dst] = r.Cc;
    }
    if (d.has_bodies) d.cdf_aff[cur][dst] = r.cdf_aff;
    if (d.has_plastic) d.plastic[cur][dst] = r.plastic;
}

__global__ void k_immigrate_commit(DeviceData d, const void* buf, uint32_t cap) {
    const uint32_t count = min(((const ShardHeader*)buf)->count, cap);
    d.counters->n_live = min(d.counters->n_live + count, d.n);
}
__global__ void k_zero_shard_counters(Counters* c) {
    c->send_count[0] = c->send_count[1] = 0;
    c->halo_count[0] = c->halo_count[1] = 0;
}
// Publishes the record counts in the buffer headers (which = 0: migration, 1: halo).
__global__ void k_write_headers(Counters* c, void* left, void* right, uint32_t cap, int which) {
    const uint32_t* cnt = which ? c->halo_count : c->send_count;
    ((ShardHeader*)left)->count = min(cnt[0], cap);
    ((ShardHeader*)right)->count = min(cnt[1], cap);
    if (cnt[0] > cap || cnt[1] > cap) c->overflow = 2u;
}
// ---- peer-to-peer exchange (NVLink stores into the neighbour's memory instead of NCCL send/recv) -------------
// The pack kernels above are handed pointers INTO THE NEIGHBOUR'S receive buffers (CUDA IPC mappings), so the
// records cross NVLink as plain stores while they are produced. k_publish then stores the record counts and,
// after a system-scope fence, the substep sequence number into the neighbour's flag word; the neighbour's
// k_wait spins on its local flag before its unpack kernels run. Buffers are double-buffered by substep parity
// (a rank cannot get more than one exchange ahead of its neighbour).
// First kernel of a peer-to-peer sharded substep: drops the dead (emigrated) tail left by the previous substep,
// resets the pack counters, snapshots the live count (immigrants are appended there) and advances the sequence.
__global__ void k_shard_tick(DeviceData d) {
    Counters* c = d.counters;
    // (the sorted total comes from its own counter, not from the bins: a grid reallocation between two substeps
    // replaces the bins, and k_begin_substep has cleared them by now anyway)
    if (c->shard_seq > 0u) c->n_live = c->sorted_total;
    c->send_count[0] = c->send_count[1] = 0;
    c->halo_count[0] = c->halo_count[1] = 0;
    c->n_base = c->n_live;
    c->shard_seq += 1u;
}

__device__ __forceinline__ void shard_wait_flag(const uint32_t* flag, uint32_t seq) {
    if (threadIdx.x == 0) {
        while (*((volatile const uint32_t*)flag) < seq) __nanosleep(100);
        __threadfence_system();
    }
    __syncthreads();
}

// Peer-to-peer immigration of BOTH neighbours' records in one kernel: waits for the neighbours' flags, appends the
// -x records at n_base and the +x records right after, and publishes the new live count.
template <int D>
__global__ void __launch_bounds__(256) k_immigrate_p2p(DeviceData d, int cur, const void* from_left, const void* from_right,
                                                       const uint32_t* flag_left, const uint32_t* flag_right, uint32_t cap) {
    const uint32_t seq = d.counters->shard_seq;
    if (from_left) shard_wait_flag(flag_left, seq);
    if (from_right) shard_wait_flag(flag_right, seq);
    const uint32_t cl = from_left ? min(((const ShardHeader*)from_left)->count, cap) : 0u;
    const uint32_t cr = from_right ? min(((const ShardHeader*)from_right)->count, cap) : 0u;
    const uint32_t n0 = d.counters->n_base;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) d.counters->n_live = min(n0 + cl + cr, d.n); // nobody reads n_live in this kernel
    TL_END(d, B200MPM_KERNEL_SHARD_MIGRATE); // (the flags have arrived: the copy below is short)
    if (i >= cl + cr) return;
    const uint32_t dst = n0 + i;
    if (dst >= d.n) {
        d.counters->overflow = 3u;
        return;
    }
    const ParticleRecord r = (i < cl) ? shard_records<ParticleRecord>(from_left)[i] : shard_records<ParticleRecord>(from_right)[i - cl];
    d.pos4[cur][dst] = r.pos4;
    d.vel4[cur][dst] = r.vel4;
    d.Fa[cur][dst] = r.Fa;
    d.Ca[cur][dst] = r.Ca;
    if (D == 3) {
        d.Fb[cur][dst] = r.Fb;
        d.Cb[cur][dst] = r.Cb;
        d.Fc[cur][dst] = r.Fc;
        d.Cc[cur][dst] = r.Cc;
    }
    if (d.has_bodies) d.cdf_aff[cur][dst] = r.cdf_aff;
    if (d.has_plastic) d.plastic[cur][dst] = r.plastic;
}

// Publishes the packed counts: header of the neighbour's buffer, system-scope fence, then the neighbour's flag.
__device__ __forceinline__ void shard_publish(Counters* c, void* left, void* right, uint32_t cap, int which,
                                              uint32_t* left_flag, uint32_t* right_flag) {
    const uint32_t* cnt = which ? c->halo_count : c->send_count;
    if (cnt[0] > cap || cnt[1] > cap) c->overflow = 2u;
    const uint32_t seq = c->shard_seq;
    if (left_flag) {
        ((ShardHeader*)left)->count = min(cnt[0], cap);
        __threadfence_system();
        *((volatile uint32_t*)left_flag) = seq;
    }
    if (right_flag) {
        ((ShardHeader*)right)->count = min(cnt[1], cap);
        __threadfence_system();
        *((volatile uint32_t*)right_flag) = seq;
    }
}
__global__ void k_publish(Counters* c, void* left, void* right, uint32_t cap, int which, uint32_t* left_flag,
                          uint32_t* right_flag) {
    shard_publish(c, left, right, cap, which, left_flag, right_flag);
}

// First kernel of a peer-to-peer sharded substep, ONE CTA: the tick (k_shard_tick), the packing of the particles
// that k_g2p listed as having left the slab (a few hundred per substep - no scan over the slab's millions), and the
// publication of the counts. The records are stored straight into the neighbours' buffers over NVLink.
template <int D>
__global__ void __launch_bounds__(256) k_emigrate_listed(DeviceData d, int cur, void* left_buf, void* right_buf, uint32_t cap,
                                                         uint32_t* left_flag, uint32_t* right_flag) {
    TL_BEGIN(d, B200MPM_KERNEL_SHARD_MIGRATE);
    Counters* c = d.counters;
    if (threadIdx.x == 0) {
        if (c->shard_seq > 0u) c->n_live = c->sorted_total;
        c->send_count[0] = c->send_count[1] = 0;
        c->halo_count[0] = c->halo_count[1] = 0;
        c->halo_ticket = 0;
        c->n_base = c->n_live;
        c->shard_seq += 1u;
    }
    __syncthreads();
    const uint32_t listed = c->emig_count;
    if (listed > d.emig_cap && threadIdx.x == 0) c->overflow = 2u; // (the unlisted ones are listed again by the next k_g2p)
    const uint32_t n = min(listed, d.emig_cap);
    ParticleRecord* left = shard_records<ParticleRecord>(left_buf);
    ParticleRecord* right = shard_records<ParticleRecord>(right_buf);
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) emigrate_particle<D>(d, cur, d.emig_list[j], left, right, cap);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        c->emig_count = 0;
        shard_publish(c, left_buf, right_buf, cap, 0, left_flag, right_flag);
    }
}

__global__ void k_wait(const Counters* c, const uint32_t* from_left, const uint32_t* from_right) {
    const uint32_t seq = c->shard_seq;
    if (from_left)
        while (*((volatile const uint32_t*)from_left) < seq) __nanosleep(200);
    if (from_right)
        while (*((volatile const uint32_t*)from_right) < seq) __nanosleep(200);
    __threadfence_system();
}

// After a substep the next buffer holds the sorted live particles in [0, total) and the parked (dead) ones after:
// dropping the tail removes the emigrants.
__global__ void k_drop_dead_tail(DeviceData d) {
    TL_BEGIN(d, B200MPM_KERNEL_SHARD_END);
    d.counters->n_live = d.counters->sorted_total;
    TL_END(d, B200MPM_KERNEL_SHARD_END);
}

// Packs the node momenta of the active blocks of the shared columns (x == slab_lo -> buffer 0, x == slab_hi ->
// buffer 1). One CTA of 64 threads per block, grid-stride.
__global__ void __launch_bounds__(CELLS_PER_BLOCK) k_halo_pack(DeviceData d, void* left_buf, void* right_buf, uint32_t cap,
                                                               uint32_t* left_flag, uint32_t* right_flag, int publish) {
    __shared__ uint32_t s_slot;
    HaloBlock* left = shard_records<HaloBlock>(left_buf);
    HaloBlock* right = shard_records<HaloBlock>(right_buf);
    const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
    const int lo = d.sim->slab_lo, hi = d.sim->slab_hi;
    TL_BEGIN(d, B200MPM_KERNEL_SHARD_HALO);
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const int4 vid = d.block_vid[b];
        const int dir = (vid.x == lo) ? 0 : (vid.x == hi) ? 1 : -1;
        if (dir < 0) continue; // (uniform per CTA)
        __syncthreads();
        if (threadIdx.x == 0) s_slot = atomicAdd(&d.counters->halo_count[dir], 1u);
        __syncthreads();
        const uint32_t slot = s_slot;
        if (slot >= cap) continue;
        HaloBlock* out = (dir == 0 ? left : right) + slot;
        if (threadIdx.x == 0) out->vid = vid;
        out->node[threadIdx.x] = d.node_mv[b * CELLS_PER_BLOCK + threadIdx.x];
    }
    if (publish) { // peer-to-peer: the last CTA to finish publishes the counts and raises the neighbours' flags
        __threadfence_system(); // my stores into the neighbours' buffers, before my ticket
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&d.counters->halo_ticket, 1u) == gridDim.x - 1u) {
            __threadfence();
            shard_publish(d.counters, left_buf, right_buf, cap, 1, left_flag, right_flag);
        }
    }
}

// Adds the neighbours' partial sums to the blocks this rank also holds (the -x neighbour's blocks lie in the column
// slab_lo, the +x neighbour's in slab_hi: no node receives from both).
template <int D>
__global__ void __launch_bounds__(CELLS_PER_BLOCK) k_halo_add(DeviceData d, const void* buf_l, const void* buf_r, uint32_t cap,
                                                              const uint32_t* flag_l, const uint32_t* flag_r) {
    // peer-to-peer: the neighbours raise the flags after their stores
    if (buf_l && flag_l) shard_wait_flag(flag_l, d.counters->shard_seq);
    if (buf_r && flag_r) shard_wait_flag(flag_r, d.counters->shard_seq);
    const uint32_t cl = buf_l ? min(((const ShardHeader*)buf_l)->count, cap) : 0u;
    const uint32_t cr = buf_r ? min(((const ShardHeader*)buf_r)->count, cap) : 0u;
    for (uint32_t k = blockIdx.x; k < cl + cr; k += gridDim.x) {
        const HaloBlock& in = (k < cl) ? shard_records<HaloBlock>(buf_l)[k] : shard_records<HaloBlock>(buf_r)[k - cl];
        const int4 vid = in.vid;
        const uint32_t hid = find_block(d.hkeys, d.hvals, d.capacity - 1, pack_key<D>(vid.x, vid.y, vid.z));
        if (hid == NONE || hid >= d.capacity) continue; // no particle of this rank reads that block
        float4 a = d.node_mv[hid * CELLS_PER_BLOCK + threadIdx.x];
        const float4 b = in.node[threadIdx.x];
        a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
        d.node_mv[hid * CELLS_PER_BLOCK + threadIdx.x] = a;
    }
    TL_END(d, B200MPM_KERNEL_SHARD_HALO);
}

// Body impulses <-> dense int32[16][6] (lin xyz, ang xyz) for the all-reduce.
__global__ void k_impulses_io(DeviceData d, int* buf, int write) {
    const uint32_t id = threadIdx.x;
    if (id >= B200MPM_MAX_BODIES) return;
    BodyDev& b = d.bodies[id];
    for (int k = 0; k < 3; ++k) {
        if (write) {
            b.imp_lin[k] = buf[id * 6 + k];
            b.imp_ang[k] = buf[id * 6 + 3 + k];
        } else {
            buf[id * 6 + k] = (id < d.sim->num_bodies) ? b.imp_lin[k] : 0;
            buf[id * 6 + 3 + k] = (id < d.sim->num_bodies) ? b.imp_ang[k] : 0;
        }
    }
}

static inline int div_up(uint64_t a, uint64_t b) { return (int)((a + b - 1) / b); }

void launch_emigrate(const LaunchCfg& c, const DeviceData& d, int cur, void* left, void* right, uint32_t cap) {
    k_zero_shard_counters<<<1, 1, 0, c.stream>>>(d.counters);
    if (d.n) {
        if (c.dim == 2) k_emigrate<2><<<div_up(d.n, 256), 256, 0, c.stream>>>(d, cur, left, right, cap);
        else k_emigrate<3><<<div_up(d.n, 256), 256, 0, c.stream>>>(d, cur, left, right, cap);
    }
    k_write_headers<<<1, 1, 0, c.stream>>>(d.counters, left, right, cap, 0);
    *c.launch_counter += 3;
}
void launch_emigrate_listed(const LaunchCfg& c, const DeviceData& d, int cur, void* left, void* right, uint32_t cap,
                            uint32_t* left_flag, uint32_t* right_flag) {
    if (c.dim == 2) k_emigrate_listed<2><<<1, 256, 0, c.stream>>>(d, cur, left, right, cap, left_flag, right_flag);
    else k_emigrate_listed<3><<<1, 256, 0, c.stream>>>(d, cur, left, right, cap, left_flag, right_flag);
    ++*c.launch_counter;
}
void launch_list_emigrants(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (!d.n) return;
    if (c.dim == 2) k_list_emigrants<2><<<div_up(d.n, 256), 256, 0, c.stream>>>(d, cur);
    else k_list_emigrants<3><<<div_up(d.n, 256), 256, 0, c.stream>>>(d, cur);
    ++*c.launch_counter;
}
void launch_shard_tick(const LaunchCfg& c, const DeviceData& d) {
    k_shard_tick<<<1, 1, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_immigrate_p2p(const LaunchCfg& c, const DeviceData& d, int cur, const void* from_left, const void* from_right,
                          const uint32_t* flag_left, const uint32_t* flag_right, uint32_t cap) {
    if (!from_left && !from_right) return;
    const uint32_t threads = cap * 2;
    if (c.dim == 2) k_immigrate_p2p<2><<<div_up(threads, 256), 256, 0, c.stream>>>(d, cur, from_left, from_right, flag_left, flag_right, cap);
    else k_immigrate_p2p<3><<<div_up(threads, 256), 256, 0, c.stream>>>(d, cur, from_left, from_right, flag_left, flag_right, cap);
    ++*c.launch_counter;
}
void launch_immigrate(const LaunchCfg& c, const DeviceData& d, int cur, const void* in, uint32_t cap) {
    if (cap == 0) return;
    if (c.dim == 2) k_immigrate<2><<<div_up(cap, 256), 256, 0, c.stream>>>(d, cur, in, cap);
    else k_immigrate<3><<<div_up(cap, 256), 256, 0, c.stream>>>(d, cur, in, cap);
    k_immigrate_commit<<<1, 1, 0, c.stream>>>(d, in, cap);
    *c.launch_counter += 2;
}
void launch_drop_dead_tail(const LaunchCfg& c, const DeviceData& d) {
    k_drop_dead_tail<<<1, 1, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_halo_pack(const LaunchCfg& c, const DeviceData& d, void* left, void* right, uint32_t cap, uint32_t* left_flag,
                      uint32_t* right_flag, bool p2p) {
    if (!p2p) k_zero_shard_counters<<<1, 1, 0, c.stream>>>(d.counters);
    k_halo_pack<<<c.num_sms * 8, CELLS_PER_BLOCK, 0, c.stream>>>(d, left, right, cap, left_flag, right_flag, p2p ? 1 : 0);
    if (!p2p) k_write_headers<<<1, 1, 0, c.stream>>>(d.counters, left, right, cap, 1);
    *c.launch_counter += p2p ? 1 : 3;
}
void launch_halo_add(const LaunchCfg& c, const DeviceData& d, const void* in_left, const void* in_right, uint32_t cap,
                     const uint32_t* flag_left, const uint32_t* flag_right) {
    if (cap == 0 || (!in_left && !in_right)) return;
    const uint32_t most = cap * ((in_left ? 1u : 0u) + (in_right ? 1u : 0u));
    int grid = (int)(most < (uint32_t)(c.num_sms * 8) ? most : (uint32_t)(c.num_sms * 8));
    if (c.dim == 2) k_halo_add<2><<<grid, CELLS_PER_BLOCK, 0, c.stream>>>(d, in_left, in_right, cap, flag_left, flag_right);
    else k_halo_add<3><<<grid, CELLS_PER_BLOCK, 0, c.stream>>>(d, in_left, in_right, cap, flag_left, flag_right);
    ++*c.launch_counter;
}
void launch_impulses_io(const LaunchCfg& c, const DeviceData& d, int* buf, int write) {
    k_impulses_io<<<1, 32, 0, c.stream>>>(d, buf, write);
    ++*c.launch_counter;
}

} // namespace b2
