// "grid sort" pass: sparse-grid block activation + counting sort of particles by (block, cell).
//
// Reference: WgGrid::queue_sort (src/grid/grid.rs:30-207) = reset_hmap, touch_particle_blocks,
// init_indirect_workgroups, update_block_particle_count, copy_particles_len_to_scan_value,
// WgPrefixSum, copy_scan_values_to_first_particles, reset, finalize_particles_sort
// (src/grid/grid.wgsl, src/grid/sort.wgsl, src/grid/prefix_sum.wgsl).
//
// B200 design: the sort key is (block header id, cell-in-block), i.e. one bin per
// grid cell, so that P2G can walk each cell's particles as a contiguous run (the reference
// keeps per-node linked lists for that, sort.wgsl:129-135). One histogram + one single-pass
// decoupled-look-back scan over B*64+1 bins (B is device-resident) replace the reference's
// capacity-length Blelloch scan; the rank returned by the histogram atomic makes the final
// scatter atomic-free.
#include "launch.h"

namespace b2 {

constexpr int SORT_THREADS = 256;
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

uint32_t scan_num_tiles(uint64_t len) { return (uint32_t)((len + SCAN_TILE - 1) / SCAN_TILE); }
__device__ __forceinline__ uint32_t scan_num_tiles_dev(uint32_t len) { return (len + SCAN_TILE - 1) / SCAN_TILE; }

// ---- reset_hmap (grid.wgsl:186-203) + clearing of last substep's bins ---------------------------
__global__ void __launch_bounds__(SORT_THREADS) k_clear(DeviceData d) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < d.capacity; i += stride) d.hkeys[i] = NONE;
    const uint32_t prev = min(d.counters->prev_active_blocks, d.capacity);
    const uint32_t nbins = prev * CELLS_PER_BLOCK + 1;
    for (uint32_t i = tid; i < nbins; i += stride) d.cell_start[i] = 0;
    const uint32_t ntiles = scan_num_tiles_dev(nbins);
    for (uint32_t i = tid; i < ntiles + 1; i += stride) d.scan_state[i] = 0ull;
}

// ---- insertion_index + mark_block_as_active (grid.wgsl:121-164, 323-334) -------------------------
template <int D>
__device__ __forceinline__ uint32_t insert_block(const DeviceData& d, int bx, int by, int bz) {
    const uint32_t mask = d.capacity - 1;
    const uint32_t key = pack_key<D>(bx, by, bz);
    uint32_t slot = hash_key(key) & mask;
    for (uint32_t k = 0; k <= mask; ++k) {
        uint32_t cur = *((volatile uint32_t*)(d.hkeys + slot));
        if (cur == key) return slot;
        if (cur == NONE) {
            uint32_t prev = atomicCAS(d.hkeys + slot, NONE, key);
            if (prev == NONE) {
                uint32_t hid = atomicAdd(&d.counters->num_active_blocks, 1u);
                if (hid < d.capacity) {
                    d.block_vid[hid] = make_int4(bx, by, bz, 0);
                    d.hvals[slot] = hid;
                }
                return slot;
            }
            if (prev == key) return slot;
        }
        slot = (slot + 1) & mask;
    }
    d.counters->overflow = 1u; // table full: the block is dropped (silent in the reference, grid.wgsl:126-128)
    return NONE;
}

// ---- touch_particle_blocks (sort.wgsl:26-36) ------------------------------------------------------
// One thread per particle; consecutive lanes that map to the same block (the common case: the
// buffers are already in last substep's sorted order) elect one leader that performs the
// 2^D insertions, the others only fetch the slot of their block from the leader.
template <int D>
__global__ void __launch_bounds__(SORT_THREADS) k_touch(DeviceData d, int cur) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < d.counters->n_live;
    const float h = d.sim->cell_width;
    const float inv_h = 1.0f / h;
    int bx = 0, by = 0, bz = 0;
    uint32_t cell = 0;
    uint32_t key = NONE;
    bool dead = false;
    if (active) {
        float4 p = d.pos4[cur][i];
        dead = (__float_as_uint(p.w) & FLAG_DEAD) != 0u; // emigrated (k_emigrate): parked, then dropped
        int cx = assoc_cell(p.x, h, inv_h), cy = assoc_cell(p.y, h, inv_h);
        bx = cx >> Dim<D>::LOG_BLOCK; // floor(c / BLOCK), grid.wgsl:286
        by = cy >> Dim<D>::LOG_BLOCK;
        cell = (cx & (Dim<D>::BLOCK - 1)) + (cy & (Dim<D>::BLOCK - 1)) * Dim<D>::BLOCK;
        if (D == 3) {
            int cz = assoc_cell(p.z, h, inv_h);
            bz = cz >> Dim<D>::LOG_BLOCK;
            cell += (cz & (Dim<D>::BLOCK - 1)) * Dim<D>::BLOCK * Dim<D>::BLOCK;
        }
        key = dead ? NONE : pack_key<D>(bx, by, bz);
    }
    const uint32_t lane = threadIdx.x & 31;
    uint32_t prev_key = __shfl_up_sync(0xffffffffu, key, 1);
    const bool leader = active && !dead && (lane == 0 || prev_key != key);
    const uint32_t leaders = __ballot_sync(0xffffffffu, leader);
    // For every run leader (usually one per warp) the 2^D insertions of blocks_associated_to_block
    // (grid.wgsl:300-320) are spread over 2^D lanes: one probe latency instead of 2^D in a row.
    uint32_t slot = NONE;
    for (uint32_t rem = leaders; rem; rem &= rem - 1) {
        const int L = __ffs(rem) - 1;
        const int lbx = __shfl_sync(0xffffffffu, bx, L), lby = __shfl_sync(0xffffffffu, by, L),
                  lbz = __shfl_sync(0xffffffffu, bz, L);
        uint32_t s = NONE;
        if (lane < (uint32_t)Dim<D>::NASSOC) {
            const int ox = lane & 1, oy = (lane >> 1) & 1, oz = (D == 3) ? (lane >> 2) & 1 : 0;
            s = insert_block<D>(d, lbx + ox, lby + oy, lbz + oz);
        }
        s = __shfl_sync(0xffffffffu, s, 0); // offset (0,..,0): the slot that identifies the particle's block
        if ((int)lane == L) slot = s;
    }
    const uint32_t below = leaders & (0xffffffffu >> (31 - lane));
    const int src = below ? (31 - __clz(below)) : 0;
    slot = __shfl_sync(0xffffffffu, slot, src);
    if (active) d.pkey[i] = (slot == NONE || dead) ? NONE : (slot * CELLS_PER_BLOCK + cell);
}

// ---- update_block_particle_count (sort.wgsl:89-99), one bin per cell ---------------------------------
__global__ void __launch_bounds__(SORT_THREADS) k_count(DeviceData d) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < d.counters->n_live;
    uint32_t ck = NONE;
    if (active) {
        const uint32_t pk = d.pkey[i];
        if (pk != NONE) {
            const uint32_t hid = d.hvals[pk >> 6];
            if (hid < d.capacity) ck = hid * CELLS_PER_BLOCK + (pk & 63u);
        }
    }
    // Runs of consecutive lanes in the same cell (the buffers are nearly sorted already) share one atomic.
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t prev = __shfl_up_sync(0xffffffffu, ck, 1);
    const bool leader = (lane == 0) || (prev != ck);
    const uint32_t leaders = __ballot_sync(0xffffffffu, leader);
    const uint32_t below = leaders & (0xffffffffu >> (31 - lane));
    const int L = 31 - __clz(below);
    const uint32_t above = (L == 31) ? 0u : (leaders >> (L + 1));
    const int run = above ? __ffs(above) : (32 - L);
    uint32_t base = 0;
    if (leader && ck != NONE) base = atomicAdd(d.cell_start + ck, (uint32_t)run);
    base = __shfl_sync(0xffffffffu, base, L);
    if (active) {
        d.pkey[i] = ck;
        d.rank[i] = base + (lane - (uint32_t)L);
    }
}

// ---- exclusive scan, single pass with decoupled look-back (replaces prefix_sum.wgsl) -----------------
// Same result as WgPrefixSum::eval_cpu (prefix_sum.rs:71-83): out[i] = sum_{j<i} in[j].
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(uint32_t* __restrict__ data, uint32_t len_value,
                                                       const Counters* __restrict__ counters, uint32_t capacity,
                                                       uint64_t* state, uint32_t* ticket) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_prefix;
    const uint32_t len = counters ? (min(counters->num_active_blocks, capacity) * CELLS_PER_BLOCK + 1) : len_value;
    // The grid is sized for the worst case (capacity); CTAs beyond the live tiles leave without a ticket, so
    // the tickets handed out are exactly 0 .. ntiles-1 and the same-address atomic stays cheap.
    if (blockIdx.x * SCAN_TILE >= len) return;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE;
    if (base >= len) return;
    const uint32_t t0 = base + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (t0 + j < len) ? data[t0 + j] : 0u;
        sum += v[j];
    }
    // inclusive warp scan of the per-thread sums
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        uint32_t x = s_warp[w];
        if (w < warp) warp_off += x;
        total += x;
    }
    if (warp == 0) { // decoupled look-back, 32 predecessors per round
        const uint64_t FLAG_A = 1ull << 62, FLAG_P = 2ull << 62;
        uint32_t running = 0;
        if (tile == 0) {
            if (lane == 0) atomicExch((unsigned long long*)(state + 0), FLAG_P | (uint64_t)total);
        } else {
            if (lane == 0) atomicExch((unsigned long long*)(state + tile), FLAG_A | (uint64_t)total);
            int j0 = (int)tile - 1; // lane l looks at tile j0 - l
            while (true) {
                const int j = j0 - (int)lane;
                uint64_t s = (j >= 0) ? *((volatile uint64_t*)(state + j)) : FLAG_P; // virtual tile -1: prefix 0
                uint64_t flag = s >> 62;
                // wait until every inspected tile has published something
                if (__any_sync(0xffffffffu, flag == 0)) continue;
                const uint32_t pmask = __ballot_sync(0xffffffffu, flag == 2);
                const int stop = pmask ? (__ffs(pmask) - 1) : 32; // nearest tile with an inclusive prefix
                uint32_t v = ((int)lane <= stop) ? (uint32_t)s : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                running += v;
                if (pmask) break;
                j0 -= 32;
            }
            if (lane == 0) atomicExch((unsigned long long*)(state + tile), FLAG_P | (uint64_t)(running + total));
        }
        if (lane == 0) s_prefix = running;
    }
    __syncthreads();
    uint32_t excl = s_prefix + warp_off + (inc - sum);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (t0 + j < len) data[t0 + j] = excl;
        excl += v[j];
    }
}

// ---- per-block preparation: neighbour table + reset (grid.wgsl:362-379) + grid_update_cdf ----------
}
#include "collide.cuh"
namespace b2 {

template <int D>
__global__ void __launch_bounds__(CELLS_PER_BLOCK) k_block_prepare(DeviceData d) {
    const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
    const uint32_t t = threadIdx.x;
    const float h = d.sim->cell_width;
    const uint32_t num_bodies = d.sim->num_bodies;
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const int4 vid = d.block_vid[b];
        if (t < Dim<D>::NASSOC) {
            int ox = t & 1, oy = (t >> 1) & 1, oz = (D == 3) ? (t >> 2) & 1 : 0;
            uint32_t key = pack_key<D>(vid.x + ox, vid.y + oy, vid.z + oz);
            d.nbr[b * Dim<D>::NASSOC + t] = (t == 0) ? b : find_block(d.hkeys, d.hvals, d.capacity - 1, key);
        }
        d.node_mv[b * CELLS_PER_BLOCK + t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d.has_bodies) { // grid_update_cdf.wgsl:16-39
            int lx = t & (Dim<D>::BLOCK - 1), ly = (t >> Dim<D>::LOG_BLOCK) & (Dim<D>::BLOCK - 1),
                lz = (D == 3) ? (t >> (2 * Dim<D>::LOG_BLOCK)) : 0;
            float pt[3] = {(float)(vid.x * Dim<D>::BLOCK + lx) * h, (float)(vid.y * Dim<D>::BLOCK + ly) * h,
                           (float)(vid.z * Dim<D>::BLOCK + lz) * h};
            NodeCdf c = collide<D>(d.bodies, num_bodies, h, pt);
            d.node_cdf[b * CELLS_PER_BLOCK + t] = make_uint4(__float_as_uint(c.distance), c.affinities, c.closest_id, 0u);
            const int any = __syncthreads_or(c.affinities != 0u);
            if (t == 0) d.block_f0[b] = any ? 1 : 0;
        }
    }
}

// ---- finalize_particles_sort (sort.wgsl:117-137): atomic-free scatter ------------------------------
template <int D>
__global__ void __launch_bounds__(SORT_THREADS) k_scatter(DeviceData d, int cur) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) d.counters->prev_active_blocks = min(d.counters->num_active_blocks, d.capacity); // for the next clear
    if (d.has_bodies) {
        // CPIC work list: a block runs the collider-aware paths iff one of the 2^D blocks its tile overlaps
        // has a node near / inside a collider. (i doubles as a block index here.)
        const uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
        if (i < nb) {
            int flag = 0;
#pragma unroll
            for (int o = 0; o < Dim<D>::NASSOC; ++o) {
                uint32_t hn = d.nbr[i * Dim<D>::NASSOC + o];
                if (hn != NONE) flag |= d.block_f0[hn];
            }
            d.block_flags[i] = (uint32_t)flag;
            if (flag && d.cell_start[i * CELLS_PER_BLOCK] != d.cell_start[(i + 1) * CELLS_PER_BLOCK])
                d.cpic_list[atomicAdd(&d.counters->num_cpic_blocks, 1u)] = i;
        }
    }
    if (i >= d.counters->n_live) return;
    uint32_t ck = d.pkey[i];
    if (ck != NONE) {
        const uint32_t dest = d.cell_start[ck] + d.rank[i];
        d.sorted_ids[dest] = i;
        if (d.has_bodies) d.cdf_aff[cur ^ 1][dest] = 0u; // default cdf; k_g2p_cdf overwrites the flagged blocks
    } else {
        // Particle of a dropped block (capacity overflow): parked after the sorted range so that
        // its state survives the ping-pong (see k_g2p tail).
        uint32_t nb = min(d.counters->num_active_blocks, d.capacity);
        uint32_t total = d.cell_start[nb * CELLS_PER_BLOCK];
        uint32_t k = atomicAdd(&d.counters->dropped_particles, 1u);
        d.sorted_ids[total + k] = i;
    }
}

// ---- launch wrappers ------------------------------------------------------------------------------------
static inline int div_up(uint64_t a, uint64_t b) { return (int)((a + b - 1) / b); }

void launch_clear(const LaunchCfg& c, const DeviceData& d) {
    k_clear<<<c.num_sms * 4, SORT_THREADS, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_touch(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0) return;
    if (c.dim == 2) k_touch<2><<<div_up(d.n, SORT_THREADS), SORT_THREADS, 0, c.stream>>>(d, cur);
    else k_touch<3><<<div_up(d.n, SORT_THREADS), SORT_THREADS, 0, c.stream>>>(d, cur);
    ++*c.launch_counter;
}
void launch_count(const LaunchCfg& c, const DeviceData& d) {
    if (d.n == 0) return;
    k_count<<<div_up(d.n, SORT_THREADS), SORT_THREADS, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_scan_cells(const LaunchCfg& c, const DeviceData& d) {
    // Upper bound on tiles: every particle activates at most 2^D blocks, and never more than capacity.
    uint64_t max_blocks = (uint64_t)d.n * (c.dim == 2 ? 4 : 8);
    if (max_blocks > d.capacity) max_blocks = d.capacity;
    uint32_t tiles = scan_num_tiles(max_blocks * CELLS_PER_BLOCK + 1);
    k_scan<<<tiles, SCAN_THREADS, 0, c.stream>>>(d.cell_start, 0u, d.counters, d.capacity, d.scan_state,
                                                  &d.counters->scan_ticket);
    ++*c.launch_counter;
}
void launch_block_prepare(const LaunchCfg& c, const DeviceData& d) {
    int grid = c.num_sms * 16;
    if (c.dim == 2) k_block_prepare<2><<<grid, CELLS_PER_BLOCK, 0, c.stream>>>(d);
    else k_block_prepare<3><<<grid, CELLS_PER_BLOCK, 0, c.stream>>>(d);
    ++*c.launch_counter;
}
void launch_scatter(const LaunchCfg& c, const DeviceData& d, int cur) {
    if (d.n == 0) return;
    // (covers both the particles and, for the CPIC work list, the active blocks: blocks <= 2^D * particles,
    // and in practice far fewer; the grid is sized for whichever is larger)
    uint64_t max_blocks = (uint64_t)d.n * (c.dim == 2 ? 4 : 8);
    if (max_blocks > d.capacity) max_blocks = d.capacity;
    uint64_t threads = d.has_bodies ? (max_blocks > d.n ? max_blocks : d.n) : d.n;
    if (c.dim == 2) k_scatter<2><<<div_up(threads, SORT_THREADS), SORT_THREADS, 0, c.stream>>>(d, cur);
    else k_scatter<3><<<div_up(threads, SORT_THREADS), SORT_THREADS, 0, c.stream>>>(d, cur);
    ++*c.launch_counter;
}
void launch_exclusive_scan_u32(const LaunchCfg& c, uint32_t* data, uint32_t len, uint64_t* scan_state,
                               uint32_t* ticket) {
    if (len == 0) return;
    uint32_t tiles = scan_num_tiles(len);
    cudaMemsetAsync(scan_state, 0, sizeof(uint64_t) * (tiles + 1), c.stream);
    cudaMemsetAsync(ticket, 0, sizeof(uint32_t), c.stream);
    k_scan<<<tiles, SCAN_THREADS, 0, c.stream>>>(data, len, nullptr, 0u, scan_state, ticket);
    ++*c.launch_counter;
}

} // namespace b2
