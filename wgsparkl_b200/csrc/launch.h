// Host-callable launch wrappers, one per kernel. Implemented in sort.cu / p2g.cu / g2p.cu / misc.cu.
#pragma once

#include "common.cuh"

namespace b2 {

// Programmatic dependent launch for the kernels of the substep's critical path (touch -> count -> scan -> scatter ->
// p2g -> g2p -> integrate): the dependent kernel is launched while its predecessor drains and parks at
// pdl_wait() (griddepcontrol.wait: full completion + memory flush of the predecessor). Every such kernel starts with
// pdl_start(). OFF by default (B200MPM_PDL=1 switches the launch attribute on; without it the device-side
// instructions are no-ops): measured, it gains 1 % on the 2M-particle dam slab and LOSES 5 % on the 1M cube - the
// kernels of this path are persistent grids that fill the machine, so an early-launched successor mostly waits for
// SM resources, and the launch gaps it could hide are ~1 us (tools/timeline.py).
bool pdl_enabled();
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#if defined(__CUDACC__)
// wait, THEN let the dependent launch: at most one kernel is ever parked, and when it starts its predecessor has passed
// its own wait - so code a kernel runs BEFORE pdl_wait() may read anything produced two or more kernels back (k_g2p
// starts its descriptor / id / record requests while k_p2g drains; only the node tile and the colours wait).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_start() {
    pdl_wait();
    pdl_trigger();
}
#endif

struct LaunchCfg {
    int dim;
    int num_sms;
    cudaStream_t stream;
    uint64_t* launch_counter; // incremented once per kernel launch (b200mpm_pipeline_launch_count)
};

// reset_hmap + bin clearing: leaves the sparse grid ready for the next k_touch (runs beside k_g2p, see api.cu).
void launch_begin_substep(const LaunchCfg& c, const DeviceData& d);
// update_world_mass_properties for all bodies (after data creation; pose / velocity writers refresh on their own).
void launch_refresh_bodies(const LaunchCfg& c, const DeviceData& d);
// "grid sort" pass (WgGrid::queue_sort, src/grid/grid.rs:30-207).
void launch_touch(const LaunchCfg& c, const DeviceData& d, int cur, bool integrate_first = false);
void launch_block_prepare(const LaunchCfg& c, const DeviceData& d); // + "grid_update_cdf" pass
void launch_scatter(const LaunchCfg& c, const DeviceData& d, int cur);
// "p2g" pass.
// "g2p_cdf" + "p2g" passes (the particle colouring runs inside the P2G kernel, p2g.cu).
void launch_p2g(const LaunchCfg& c, const DeviceData& d, int cur);
// "grid_update" + "g2p" + "particles_update" passes (fused).
void launch_g2p_update(const LaunchCfg& c, const DeviceData& d, int cur);
// "integrate_bodies" pass.
void launch_integrate_bodies(const LaunchCfg& c, const DeviceData& d);

// Stand-alone exclusive scan of a device vector (b200mpm_prefix_sum_u32).
void launch_exclusive_scan_u32(const LaunchCfg& c, uint32_t* data, uint32_t len, uint64_t* scan_state,
                               uint32_t* ticket);
uint32_t scan_num_tiles(uint64_t len);

// Readback helpers (device -> staging in caller order).
void launch_gather_positions(const LaunchCfg& c, const DeviceData& d, int cur, float4* out, int unordered = 0);
void launch_gather_particles(const LaunchCfg& c, const DeviceData& d, int cur, b200mpm_particle* out, uint32_t* ids);

// Slab sharding (shard.cu)
void launch_emigrate(const LaunchCfg& c, const DeviceData& d, int cur, void* left, void* right, uint32_t cap);
// peer-to-peer: tick + the particles k_g2p listed + publication, one CTA
void launch_emigrate_listed(const LaunchCfg& c, const DeviceData& d, int cur, void* left, void* right, uint32_t cap,
                            uint32_t* left_flag, uint32_t* right_flag);
void launch_list_emigrants(const LaunchCfg& c, const DeviceData& d, int cur);
void launch_shard_tick(const LaunchCfg& c, const DeviceData& d);
void launch_immigrate_p2p(const LaunchCfg& c, const DeviceData& d, int cur, const void* from_left, const void* from_right,
                          const uint32_t* flag_left, const uint32_t* flag_right, uint32_t cap);
void launch_immigrate(const LaunchCfg& c, const DeviceData& d, int cur, const void* in, uint32_t cap);
void launch_drop_dead_tail(const LaunchCfg& c, const DeviceData& d);
void launch_halo_pack(const LaunchCfg& c, const DeviceData& d, void* left, void* right, uint32_t cap,
                      uint32_t* left_flag = nullptr, uint32_t* right_flag = nullptr, bool p2p = false);
void launch_halo_add(const LaunchCfg& c, const DeviceData& d, const void* in_left, const void* in_right, uint32_t cap,
                     const uint32_t* flag_left = nullptr, const uint32_t* flag_right = nullptr);
void launch_impulses_io(const LaunchCfg& c, const DeviceData& d, int* buf, int write);
void launch_gather_grid(const LaunchCfg& c, const DeviceData& d, b200mpm_block_info* blocks, b200mpm_node* nodes,
                        uint32_t max_blocks);
void launch_gather_sorted_ids(const LaunchCfg& c, const DeviceData& d, int cur, int indirect, uint32_t* out);
void launch_transform_rigid(const LaunchCfg& c, const DeviceData& d);
void launch_touch_rigid(const LaunchCfg& c, const DeviceData& d);
void launch_p2g_cdf(const LaunchCfg& c, const DeviceData& d);
void launch_prep_vertex_buffer(const LaunchCfg& c, const DeviceData& d, int cur, b200mpm_instance* inst, uint32_t mode);
void launch_write_poses(const LaunchCfg& c, const DeviceData& d, const b200mpm_pose* poses, uint32_t n);
void launch_write_vels(const LaunchCfg& c, const DeviceData& d, const b200mpm_velocity* vels, uint32_t n);
void launch_read_poses(const LaunchCfg& c, const DeviceData& d, b200mpm_pose* poses, b200mpm_velocity* vels, uint32_t n);

} // namespace b2
