// sc_host.cu -- host-side plumbing of libsinglecarrier_b200.so that is not modem arithmetic:
//   * page-locked host memory on the GPU's NUMA node and in-place registration of caller memory,
//   * the host<->device copy probe that establishes the platform (PCIe / host memory) ceiling the
//     host entry point sc_rx_frames_host is judged against,
//   * the NCCL bridge for the path's only collective: one all-reduce of the lock / bit counters
//     (SURVEY section 8e).  NCCL is bound at run time (dlopen) so the library loads without it.
#include <dlfcn.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <mutex>
#include <string>
#include <vector>

#include "sc_common.cuh"
#include "sc_kernels.h"

namespace sc {
int api_fail(int code, const char *msg);
}
using namespace sc;

#define HCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[256];                                                                        \
            snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return api_fail(e_ == cudaErrorMemoryAllocation ? SC_ENOMEM : SC_ECUDA, buf_);         \
        }                                                                                          \
    } while (0)

// ---- NUMA placement ------------------------------------------------------------------------------
static int read_int_file(const char *path, int fallback) {
    FILE *f = fopen(path, "r");
    if (!f) return fallback;
    int v = fallback;
    if (fscanf(f, "%d", &v) != 1) v = fallback;
    fclose(f);
    return v;
}

// "0-15,32-47" -> cpu set
static bool parse_cpulist(const char *path, cpu_set_t *set) {
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char line[4096];
    bool any = false;
    CPU_ZERO(set);
    if (fgets(line, sizeof line, f)) {
        for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            const int k = sscanf(tok, "%d-%d", &a, &b);
            if (k == 1) b = a;
            if (k >= 1)
                for (int c = a; c <= b && c < CPU_SETSIZE; c++) {
                    CPU_SET(c, set);
                    any = true;
                }
        }
    }
    fclose(f);
    return any;
}

static int device_numa_node(int device) {
    char bus[64] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bus; *c; c++) *c = (char) tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    return read_int_file(path.c_str(), -1);
}

extern "C" int sc_host_alloc(void **ptr, size_t bytes, int device, int *numa_node) {
    if (!ptr || bytes == 0) return api_fail(SC_EINVAL, "sc_host_alloc: bad arguments");
    *ptr = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return api_fail(SC_ECUDA, "sc_host_alloc: no CUDA device");
    }
    if (device < 0 || device >= ndev) return api_fail(SC_EINVAL, "sc_host_alloc: bad device");
    HCU(cudaSetDevice(device));
    const int node = device_numa_node(device);
    if (numa_node) *numa_node = node;
    // first-touch placement: run on the node's CPUs while the driver allocates and we touch the pages
    cpu_set_t old_set, node_set;
    bool narrowed = false;
    if (node >= 0 && sched_getaffinity(0, sizeof old_set, &old_set) == 0) {
        char path[128];
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        if (parse_cpulist(path, &node_set)) {
            cpu_set_t both;
            CPU_AND(&both, &node_set, &old_set);
            if (CPU_COUNT(&both) > 0 && sched_setaffinity(0, sizeof both, &both) == 0) narrowed = true;
        }
    }
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) {
        volatile char *p = (volatile char *) *ptr;
        for (size_t off = 0; off < bytes; off += 4096) p[off] = 0;
    }
    if (narrowed) sched_setaffinity(0, sizeof old_set, &old_set);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return api_fail(e == cudaErrorMemoryAllocation ? SC_ENOMEM : SC_ECUDA, "sc_host_alloc: cudaHostAlloc failed");
    }
    return SC_OK;
}

extern "C" int sc_host_free(void *ptr) {
    if (!ptr) return SC_OK;
    HCU(cudaFreeHost(ptr));
    return SC_OK;
}

extern "C" int sc_host_register(void *ptr, size_t bytes) {
    if (!ptr || bytes == 0) return api_fail(SC_EINVAL, "sc_host_register: bad arguments");
    HCU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return SC_OK;
}

extern "C" int sc_host_unregister(void *ptr) {
    if (!ptr) return api_fail(SC_EINVAL, "sc_host_unregister: null pointer");
    HCU(cudaHostUnregister(ptr));
    return SC_OK;
}

// ---- copy probe ----------------------------------------------------------------------------------
static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

extern "C" int sc_h2d_probe(int device, size_t buffer_bytes, size_t row_bytes, size_t src_pitch_bytes, double min_seconds,
                            int d2h, double *gbytes_per_s) {
    if (!gbytes_per_s || buffer_bytes < 4096 || (row_bytes && (src_pitch_bytes < row_bytes || row_bytes > buffer_bytes)))
        return api_fail(SC_EINVAL, "sc_h2d_probe: bad arguments");
    *gbytes_per_s = 0.0;
    void *host = nullptr;
    int rc = sc_host_alloc(&host, buffer_bytes, device, nullptr);
    if (rc != SC_OK) return rc;
    struct Guard {
        void *h, *d = nullptr;
        cudaStream_t st = nullptr;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        ~Guard() {
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            if (st) cudaStreamDestroy(st);
            if (d) cudaFree(d);
            if (h) cudaFreeHost(h);
        }
    } g{host};
    HCU(cudaMalloc(&g.d, buffer_bytes));
    HCU(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
    HCU(cudaEventCreate(&g.e0));
    HCU(cudaEventCreate(&g.e1));
    const size_t rows = row_bytes ? std::max<size_t>(1, (buffer_bytes - row_bytes) / src_pitch_bytes + 1) : 0;
    const size_t moved = row_bytes ? rows * row_bytes : buffer_bytes;
    auto copy = [&]() -> cudaError_t {
        if (!row_bytes)
            return d2h ? cudaMemcpyAsync(g.h, g.d, buffer_bytes, cudaMemcpyDeviceToHost, g.st)
                       : cudaMemcpyAsync(g.d, g.h, buffer_bytes, cudaMemcpyHostToDevice, g.st);
        // device side dense (pitch = row), host side strided: the staging pattern of sc_rx_frames_host
        return d2h ? cudaMemcpy2DAsync(g.h, src_pitch_bytes, g.d, row_bytes, row_bytes, rows, cudaMemcpyDeviceToHost, g.st)
                   : cudaMemcpy2DAsync(g.d, row_bytes, g.h, src_pitch_bytes, row_bytes, rows, cudaMemcpyHostToDevice, g.st);
    };
    HCU(copy());                                           // warm-up
    HCU(cudaStreamSynchronize(g.st));
    double total_ms = 0.0, total_bytes = 0.0;
    const double t_end = now_s() + std::max(min_seconds, 0.0);
    do {
        HCU(cudaEventRecord(g.e0, g.st));
        for (int k = 0; k < 4; k++) HCU(copy());
        HCU(cudaEventRecord(g.e1, g.st));
        HCU(cudaStreamSynchronize(g.st));
        float ms = 0.f;
        HCU(cudaEventElapsedTime(&ms, g.e0, g.e1));
        total_ms += ms;
        total_bytes += 4.0 * (double) moved;
    } while (now_s() < t_end);
    *gbytes_per_s = total_bytes / (total_ms * 1e-3) / 1e9;
    return SC_OK;
}

// ---- NCCL bridge ---------------------------------------------------------------------------------
// Only the handful of entry points the reduction needs, declared here so that neither nccl.h nor a
// link-time dependency is required (the ABI of these functions is stable across NCCL 2.x).
namespace {
typedef struct { char internal[SC_NCCL_UNIQUE_ID_BYTES]; } nccl_unique_id;
typedef int (*fn_get_unique_id)(nccl_unique_id *);
typedef int (*fn_comm_init_rank)(void **, int, nccl_unique_id, int);
typedef int (*fn_comm_init_all)(void **, int, const int *);
typedef int (*fn_comm_destroy)(void *);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_get_error_string)(int);
typedef int (*fn_group)(void);
const int NCCL_UINT64 = 5, NCCL_SUM = 0;               // ncclDataType_t / ncclRedOp_t values, nccl.h

struct Nccl {
    void *lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_init_all comm_init_all = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_get_error_string error_string = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    std::once_flag once;
    std::string why;
} g_nccl;

void nccl_bind() {
    // prefer the copy that is already in the process (PyTorch loads its own libnccl.so.2): two NCCL
    // instances would work but double the bootstrap
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        g_nccl.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
        return;
    }
    g_nccl.lib = h;
    g_nccl.get_unique_id = (fn_get_unique_id) dlsym(h, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank) dlsym(h, "ncclCommInitRank");
    g_nccl.comm_init_all = (fn_comm_init_all) dlsym(h, "ncclCommInitAll");
    g_nccl.comm_destroy = (fn_comm_destroy) dlsym(h, "ncclCommDestroy");
    g_nccl.all_reduce = (fn_all_reduce) dlsym(h, "ncclAllReduce");
    g_nccl.error_string = (fn_get_error_string) dlsym(h, "ncclGetErrorString");
    g_nccl.group_start = (fn_group) dlsym(h, "ncclGroupStart");
    g_nccl.group_end = (fn_group) dlsym(h, "ncclGroupEnd");
    if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.comm_init_all || !g_nccl.comm_destroy ||
        !g_nccl.all_reduce || !g_nccl.group_start || !g_nccl.group_end) {
        g_nccl.why = "libnccl.so.2 lacks an expected symbol";
        g_nccl.lib = nullptr;
    }
}

int nccl_ready() {
    std::call_once(g_nccl.once, nccl_bind);
    if (!g_nccl.lib) return api_fail(SC_ESTATE, ("NCCL unavailable: " + g_nccl.why).c_str());
    return SC_OK;
}

int nccl_check(int r, const char *what) {
    if (r == 0) return SC_OK;
    char buf[256];
    snprintf(buf, sizeof buf, "%s: NCCL error %d (%s)", what, r, g_nccl.error_string ? g_nccl.error_string(r) : "?");
    return api_fail(SC_ECUDA, buf);
}
}  // namespace

extern "C" int sc_comm_unique_id(void *id128) {
    if (!id128) return api_fail(SC_EINVAL, "sc_comm_unique_id: null pointer");
    int rc = nccl_ready();
    if (rc != SC_OK) return rc;
    nccl_unique_id id;
    if ((rc = nccl_check(g_nccl.get_unique_id(&id), "ncclGetUniqueId")) != SC_OK) return rc;
    memcpy(id128, &id, sizeof id);
    return SC_OK;
}

extern "C" int sc_comm_init_rank(void **comm, int n_ranks, int rank, const void *id128, int device) {
    if (!comm || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return api_fail(SC_EINVAL, "sc_comm_init_rank: bad arguments");
    int rc = nccl_ready();
    if (rc != SC_OK) return rc;
    HCU(cudaSetDevice(device));
    nccl_unique_id id;
    memcpy(&id, id128, sizeof id);
    return nccl_check(g_nccl.comm_init_rank(comm, n_ranks, id, rank), "ncclCommInitRank");
}

extern "C" int sc_comm_init_all(void **comms, int n_devices, const int *devices) {
    if (!comms || n_devices < 1) return api_fail(SC_EINVAL, "sc_comm_init_all: bad arguments");
    int rc = nccl_ready();
    if (rc != SC_OK) return rc;
    return nccl_check(g_nccl.comm_init_all(comms, n_devices, devices), "ncclCommInitAll");
}

extern "C" int sc_comm_destroy(void *comm) {
    if (!comm) return SC_OK;
    int rc = nccl_ready();
    if (rc != SC_OK) return rc;
    return nccl_check(g_nccl.comm_destroy(comm), "ncclCommDestroy");
}

extern "C" int sc_reduce_stats(uint64_t *counters, int n_counters, void *nccl_comm, void *stream) {
    if (!counters || n_counters < 1 || !nccl_comm) return api_fail(SC_EINVAL, "sc_reduce_stats: bad arguments");
    int rc = nccl_ready();
    if (rc != SC_OK) return rc;
    return nccl_check(g_nccl.all_reduce(counters, counters, (size_t) n_counters, NCCL_UINT64, NCCL_SUM, nccl_comm,
                                        (cudaStream_t) stream), "ncclAllReduce");
}

// one process driving several ranks brackets their sc_reduce_stats calls with these (ncclGroupStart/End)
extern "C" int sc_comm_group_start(void) {
    int rc = nccl_ready();
    if (rc != SC_OK) return rc;
    return nccl_check(g_nccl.group_start(), "ncclGroupStart");
}
extern "C" int sc_comm_group_end(void) {
    int rc = nccl_ready();
    if (rc != SC_OK) return rc;
    return nccl_check(g_nccl.group_end(), "ncclGroupEnd");
}

// ---- device memory for callers without a CUDA binding (plain C, Go, Rust ...) -------------------------
extern "C" int sc_device_malloc(int device, size_t bytes, void **ptr) {
    if (!ptr || bytes == 0) return api_fail(SC_EINVAL, "sc_device_malloc: bad arguments");
    *ptr = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return api_fail(SC_ECUDA, "sc_device_malloc: no CUDA device");
    }
    if (device < 0 || device >= ndev) return api_fail(SC_EINVAL, "sc_device_malloc: bad device");
    HCU(cudaSetDevice(device));
    HCU(cudaMalloc(ptr, bytes));
    HCU(cudaMemset(*ptr, 0, bytes));
    return SC_OK;
}
extern "C" int sc_device_free(int device, void *ptr) {
    if (!ptr) return SC_OK;
    HCU(cudaSetDevice(device));
    HCU(cudaFree(ptr));
    return SC_OK;
}
extern "C" int sc_device_copy(int device, void *dst, const void *src, size_t bytes, int kind) {
    if (!dst || !src || kind < 0 || kind > 2) return api_fail(SC_EINVAL, "sc_device_copy: bad arguments");
    HCU(cudaSetDevice(device));
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    HCU(cudaMemcpy(dst, src, bytes, k));
    return SC_OK;
}
extern "C" int sc_device_synchronize(int device) {
    HCU(cudaSetDevice(device));
    HCU(cudaDeviceSynchronize());
    return SC_OK;
}
