"""The consumer path of b200mpm_prep_vertex_buffer (src_testbed/prep_vertex_buffer.rs:73-111): the renderer owns the
instance buffer, exports it as an opaque POSIX file descriptor (Vulkan: VK_KHR_external_memory_fd), and the compute
side imports it with cudaImportExternalMemory and writes into the mapped pointer. There is no Vulkan driver in the
image, so the exporting side is played by the CUDA driver's own shareable allocation
(cuMemCreate + cuMemExportToShareableHandle(POSIX_FILE_DESCRIPTOR)); the import / map / write / read-back path is the
one a wgpu-hal host would use."""
import numpy as np
import pytest

from wgsparkl_b200 import abi, scenes
from wgsparkl_b200.pipeline import MpmData

pytestmark = pytest.mark.gpu


def _ok(res, what):
    err = res[0]
    assert int(err) == 0, "%s failed: %s" % (what, err)
    return res[1] if len(res) == 2 else res[1:]


def test_prep_vertex_buffer_into_imported_external_memory(pipe3):
    import torch
    from cuda.bindings import driver, runtime

    torch.cuda.init()
    torch.zeros(1, device="cuda")  # primary context current
    scene = scenes.elastic_cube_3d(12, y_offset=3.0)
    data = MpmData(pipe3, scene["params"], scene["particles"], scene["bodies"], scene["cell_width"], scene["grid_capacity"])
    pipe3.queue_step(data, 3)
    n = len(scene["particles"])
    want_bytes = n * abi.instance_dtype.itemsize

    # --- the "renderer": a device allocation that can be exported as a file descriptor
    prop = driver.CUmemAllocationProp()
    prop.type = driver.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = driver.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    prop.requestedHandleTypes = driver.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = _ok(driver.cuMemGetAllocationGranularity(prop, driver.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM), "granularity")
    size = (want_bytes + gran - 1) // gran * gran
    handle = _ok(driver.cuMemCreate(size, prop, 0), "cuMemCreate")
    fd = _ok(driver.cuMemExportToShareableHandle(handle, driver.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export")
    # its own mapping, to initialise the buffer and to look at the result
    own = _ok(driver.cuMemAddressReserve(size, 0, 0, 0), "reserve")
    _ok(driver.cuMemMap(own, size, 0, handle, 0), "map")
    acc = driver.CUmemAccessDesc()
    acc.location.type = driver.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = 0
    acc.flags = driver.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    _ok(driver.cuMemSetAccess(own, size, [acc], 1), "access")
    init = np.zeros(n, dtype=abi.instance_dtype)
    init["base_color"] = 0.25
    init["position"][:, 3] = 9.0
    _ok(runtime.cudaMemcpy(int(own), init.ctypes.data, want_bytes, runtime.cudaMemcpyKind.cudaMemcpyHostToDevice), "upload")

    # --- the compute side: import the fd, map it, hand the pointer to the C ABI
    desc = runtime.cudaExternalMemoryHandleDesc()
    desc.type = runtime.cudaExternalMemoryHandleType.cudaExternalMemoryHandleTypeOpaqueFd
    desc.handle.fd = int(fd)
    desc.size = size
    res = runtime.cudaImportExternalMemory(desc)
    if int(res[0]) != 0:
        pytest.skip("cudaImportExternalMemory does not accept this driver's own exported fd: %s" % res[0])
    ext = res[1]
    bdesc = runtime.cudaExternalMemoryBufferDesc()
    bdesc.offset = 0
    bdesc.size = size
    bdesc.flags = 0
    imported = _ok(runtime.cudaExternalMemoryGetMappedBuffer(ext, bdesc), "mapped buffer")
    assert int(imported) != int(own)
    pipe3.prep_vertex_buffer(data, int(imported), abi.RENDER_DEFAULT)
    pipe3.sync()

    got = np.zeros(n, dtype=abi.instance_dtype)
    _ok(runtime.cudaMemcpy(got.ctypes.data, int(own), want_bytes, runtime.cudaMemcpyKind.cudaMemcpyDeviceToHost), "download")
    ref = data.read_particles()
    assert np.array_equal(got["position"][:, :3], ref["position"])  # written through the imported mapping
    assert np.all(got["position"][:, 3] == 9.0) and np.all(got["base_color"] == 0.25)  # untouched lanes survive
    assert np.isfinite(got["deformation"]).all() and np.abs(got["deformation"][:, :, :3]).max() > 0.5  # F ~ identity

    _ok(runtime.cudaFree(int(imported)), "free mapped")
    _ok(runtime.cudaDestroyExternalMemory(ext), "destroy")
    _ok(driver.cuMemUnmap(own, size), "unmap")
    _ok(driver.cuMemAddressFree(own, size), "address free")
    _ok(driver.cuMemRelease(handle), "release")
    data.close()
