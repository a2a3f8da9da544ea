// probe_fabric.cu -- what the box offers for peer / multicast memory (prints one JSON object).
// Used once per box type to decide whether the NVSwitch multicast path (multimem.*) can run:
// CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, handle types, peer access matrix, P2P atomics.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>

int main()
{
    if (cuInit(0) != CUDA_SUCCESS) {
        printf("{\"error\": \"cuInit failed\"}\n");
        return 1;
    }
    int n = 0;
    cuDeviceGetCount(&n);
    printf("{\"devices\": %d, \"per_device\": [", n);
    for (int d = 0; d < n; ++d) {
        CUdevice dev;
        cuDeviceGet(&dev, d);
        int mc = -1, fabric = -1, posix = -1, vmm = -1, gdr = -1;
        cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
        cuDeviceGetAttribute(&fabric, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev);
        cuDeviceGetAttribute(&posix, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
        cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
        cuDeviceGetAttribute(&gdr, CU_DEVICE_ATTRIBUTE_GPU_DIRECT_RDMA_SUPPORTED, dev);
        size_t gran_min = 0, gran_rec = 0;
        if (mc == 1 && n >= 2) {
            CUmulticastObjectProp p = {};
            p.numDevices = (unsigned)n;
            p.size = 2u << 20;
            p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
            cuMulticastGetGranularity(&gran_min, &p, CU_MULTICAST_GRANULARITY_MINIMUM);
            cuMulticastGetGranularity(&gran_rec, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED);
        }
        int persist = 0, l2 = 0, window = 0;
        cudaDeviceGetAttribute(&persist, cudaDevAttrMaxPersistingL2CacheSize, d);
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, d);
        cudaDeviceGetAttribute(&window, cudaDevAttrMaxAccessPolicyWindowSize, d);
        printf("%s{\"device\": %d, \"multicast\": %d, \"fabric_handles\": %d, \"posix_fd_handles\": %d, \"vmm\": %d, "
               "\"mc_granularity_min\": %zu, \"mc_granularity_rec\": %zu, \"l2_bytes\": %d, \"max_persisting_l2\": %d, "
               "\"max_access_policy_window\": %d}",
               d ? ", " : "", d, mc, fabric, posix, vmm, gran_min, gran_rec, l2, persist, window);
    }
    printf("], \"peer\": [");
    for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            int can = 0, atom = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, a, b);
            printf("%s[%d, %d, %d, %d]", (a == 0 && b == 1) ? "" : ", ", a, b, can, atom);
        }
    printf("]}\n");
    return 0;
}
