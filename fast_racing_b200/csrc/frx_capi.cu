// Implementation of the C ABI declared in include/frx.h.
#include <cuda_runtime.h>

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is opened at run time (dlopen), libfrx.so does not link it
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "../../include/frx.h"
#include "corridor.hpp"
#include "corridor_kernels.cuh"
#include "device_batch.cuh"
#include "geometry_kernels.cuh"
#include "host_model.hpp"
#include "solver.cuh"
#include "waypoint_nls.hpp"

#define FRX_DECLARE_EVAL_API(NS)                                                                                                   \
    namespace NS {                                                                                                                 \
    size_t forward_smem_bytes(int M, int coarseN);                                                                                 \
    size_t backward_smem_bytes(int M, int coarseN);                                                                                \
    frx::PenaltyLaunch plan_penalty(const frx::BatchView& bv);                                                                     \
    cudaError_t configure_kernels(size_t fwd_smem, size_t bwd_smem, size_t pen_smem);                                              \
    cudaError_t launch_eval_window(const frx::BatchView& bv0, const frx::PenaltyLaunch& pl, size_t fwd_smem, size_t bwd_smem,      \
                                   const double* x_dev, double* cost_dev, double* grad_dev, cudaStream_t st, int t0, int t1,       \
                                   const int* piece_off_host);                                                                     \
    cudaError_t launch_eval_pipelined(const frx::BatchView& bv0, const frx::PenaltyLaunch& pl, size_t fwd_smem, size_t bwd_smem,   \
                                      const double* x_dev, double* cost_dev, double* grad_dev, cudaStream_t st, cudaStream_t aux,  \
                                      cudaEvent_t ev_fork, cudaEvent_t ev_k1, cudaEvent_t ev_join, const int* piece_off_host,      \
                                      int nsub);                                                                                   \
    cudaError_t profile_eval(const frx::BatchView& bv, const frx::PenaltyLaunch& pl, size_t fwd_smem, size_t bwd_smem,             \
                             const double* x_dev, double* cost_dev, double* grad_dev, cudaStream_t st, int reps, double* ms);      \
    double measure_fp64_tflops();                                                                                                  \
    cudaError_t fetch_phase_clocks(long long* out96);                                                                              \
    cudaError_t fetch_cta_spans(unsigned long long* out);                                                                          \
    }
FRX_DECLARE_EVAL_API(frx_fast)
FRX_DECLARE_EVAL_API(frx_exact)

using namespace frx;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t alloc(size_t n) {
        bytes = n ? n : 8;
        return cudaMalloc(&p, bytes);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
    }
};

struct frx_batch {
    int device = 0;
    std::vector<TrajectoryModel> models;
    std::vector<int> piece_off, x_off, coarse_off;
    int max_M = 0, max_coarse = 0;
    BatchView bv{};
    PenaltyLaunch pl{};
    size_t fwd_smem = 0, bwd_smem = 0;      // fast build
    size_t fwd_smem_x = 0, bwd_smem_x = 0;  // reference-order build
    bool exact = false;                      // frx_batch_set_reference_order
    bool last_solver_exact = false;
    cudaStream_t stream = nullptr, aux = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_k1 = nullptr, ev_join = nullptr;
    int e2e_chunks = 2;  // chunks of the host-buffer path: copies of one chunk overlap the kernels of the other (FRX_E2E_CHUNKS)
    int solver_slots = 0;  // frx_batch_set_solver_slots: 0 = one slot per trajectory
    int nsub = 1;  // sub-batches of the pipelined evaluation (FRX_EVAL_SUBBATCHES); measured on B200: K2 fills the register
                   // file (8 CTAs x 64 thr x 128 regs), so K1/K3 cannot co-reside and pipelining only adds tails -> default off
    cudaError_t eval_on(const double* x, double* cost, double* grad, cudaStream_t st) {
        if (exact)
            return frx_exact::launch_eval_pipelined(bv, pl, fwd_smem_x, bwd_smem_x, x, cost, grad, st, nullptr, nullptr, nullptr,
                                                    nullptr, piece_off.data(), 1);
        return frx_fast::launch_eval_pipelined(bv, pl, fwd_smem, bwd_smem, x, cost, grad, st, aux, ev_fork, ev_k1, ev_join,
                                               piece_off.data(), nsub);
    }
    cudaError_t eval_window_on(const double* x, double* cost, double* grad, cudaStream_t st, int t0, int t1) {
        if (exact) return frx_exact::launch_eval_window(bv, pl, fwd_smem_x, bwd_smem_x, x, cost, grad, st, t0, t1, piece_off.data());
        return frx_fast::launch_eval_window(bv, pl, fwd_smem, bwd_smem, x, cost, grad, st, t0, t1, piece_off.data());
    }
    double last_ms = 0.0;
    std::vector<DevBuf> bufs;
    // persistent device-side x / cost / grad for the host-buffer API
    double *x_dev = nullptr, *g_dev = nullptr, *cost_dev = nullptr;
    double *x_pin = nullptr, *g_pin = nullptr, *cost_pin = nullptr;
    frx_fast::SolverState* solver = nullptr;     // lazily created by frx_batch_optimize
    frx_exact::SolverState* solver_x = nullptr;  // reference-order solver
    // result of the last optimize (for argmin)
    double* final_cost_dev = nullptr;
    bool have_final = false;

    template <class T>
    cudaError_t upload(const std::vector<T>& h, const T** out) {
        DevBuf b;
        cudaError_t e = b.alloc(sizeof(T) * h.size());
        if (e != cudaSuccess) return e;
        bufs.push_back(b);
        if (!h.empty()) e = cudaMemcpy(b.p, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
        *out = (const T*)b.p;
        return e;
    }
    template <class T>
    cudaError_t scratch(size_t n, T** out) {
        DevBuf b;
        cudaError_t e = b.alloc(sizeof(T) * n);
        if (e != cudaSuccess) return e;
        bufs.push_back(b);
        e = cudaMemset(b.p, 0, b.bytes);
        *out = (T*)b.p;
        return e;
    }
};

// ---- NUMA placement of pinned host memory -------------------------------------------------------------------------
// One rank per GPU: eight ranks whose pinned x / cost / grad buffers all sit on one socket drag every copy of the far GPUs
// across the inter-socket link (round 1: end-to-end efficiency 0.54 at 8 GPUs while the device-resident value scaled at
// 0.99).  The buffers are therefore allocated with the calling thread moved to the CPUs of the GPU's own NUMA node and the
// memory policy preferring that node; both are restored afterwards.  Best effort: every failure leaves the default
// placement (no libnuma in the image: sysfs + raw syscalls).  FRX_NUMA=0 disables it.
static int device_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return -1;
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    char path[128];
    std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = std::fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}
static bool node_cpuset(int node, cpu_set_t* set) {
    char path[96];
    std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE* f = std::fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(set);
    int a = 0, b2 = 0;
    bool any = false;
    char sep = 0;
    while (std::fscanf(f, "%d", &a) == 1) {
        b2 = a;
        if (std::fscanf(f, "%c", &sep) == 1 && sep == '-') {
            if (std::fscanf(f, "%d", &b2) != 1) break;
            if (std::fscanf(f, "%c", &sep) != 1) sep = 0;
        }
        for (int c = a; c <= b2 && c < CPU_SETSIZE; c++) CPU_SET(c, set), any = true;
        if (sep != ',') break;
    }
    std::fclose(f);
    return any;
}
struct NumaScope {
    bool moved = false, policy = false;
    cpu_set_t old;
    explicit NumaScope(int device, bool keep = false) : keep_(keep) {
        const char* e = getenv("FRX_NUMA");
        if (e && e[0] == '0') return;
        node = device_numa_node(device);
        if (node < 0) return;
        cpu_set_t want, cur;
        if (!node_cpuset(node, &want) || sched_getaffinity(0, sizeof(cur), &cur) != 0) return;
        old = cur;
        CPU_AND(&want, &want, &cur);  // never leave the mask the process was given
        if (CPU_COUNT(&want) == 0) return;
        if (sched_setaffinity(0, sizeof(want), &want) == 0) moved = true;
        if (node < 64) {
            unsigned long mask = 1ul << node;
            policy = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, sizeof(mask) * 8 + 1) == 0;
        }
    }
    ~NumaScope() {
        if (keep_) return;
        if (policy) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
        if (moved) sched_setaffinity(0, sizeof(old), &old);
    }
    int node = -1;

private:
    bool keep_;
};

static int cuda_fail(cudaError_t e, const char* what) {
    std::fprintf(stderr, "[frx] CUDA error in %s: %s\n", what, cudaGetErrorString(e));
    return FRX_ERR_CUDA;
}
#define CK(call)                                         \
    do {                                                 \
        cudaError_t e_ = (call);                         \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

// Every entry point runs on its handle's device and gives the caller's current device back on return (a planner with
// several GPUs, or torch in the same process, keeps its own current device).
struct DeviceScope {
    int prev = -1;
    bool ok = true;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
        else prev = -1;  // nothing to restore
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
#define ON_DEVICE(dev)             \
    DeviceScope device_scope_(dev); \
    if (!device_scope_.ok) return cuda_fail(cudaErrorInvalidDevice, "cudaSetDevice")

// ---- batched V-polytope extraction on the device ----------------------------------------------------
// For every problem that does not bring its own V-polytopes: all 2 n_poly - 1 polytopes (P_0, P_0^P_1, P_1, ...,
// se3gcopter_cpu.hpp:1031-1074) of all problems go through ONE launch of enumerate_vertices_kernel; the result is handed to
// build_model through the "supplied V-polytopes" path in the reference's "origin + edges" form (:1049).  A polytope the
// kernel declines (GEO_FALLBACK) sends its whole problem to the host enumeration (same algorithm, same result).
// FRX_HOST_GEOMETRY=1 forces the host path for everything.
struct DeviceGeometry {
    std::vector<std::vector<int>> off;       // per problem: 2 n_poly offsets (in vertices), empty = not computed here
    std::vector<std::vector<double>> verts;  // per problem: 3 doubles per vertex column
    std::vector<int> status;                 // per problem: FRX_OK / FRX_ERR_INFEASIBLE_GEOMETRY
};
static int device_geometry(const frx_problem* problems, int count, DeviceGeometry& G) {
    G.off.assign(count, {}), G.verts.assign(count, {}), G.status.assign(count, FRX_OK);
    const char* env = getenv("FRX_HOST_GEOMETRY");
    if (env && env[0] == '1') return FRX_OK;
    std::vector<double> planes;
    std::vector<int2> polys;
    std::vector<int> vstart, first_poly(count + 1, 0);
    int slots = 0;
    for (int i = 0; i < count; i++) {
        const frx_problem& p = problems[i];
        first_poly[i] = (int)polys.size();
        if ((p.vert_off && p.verts) || p.n_poly < 1 || !p.face_off || !p.faces) continue;
        bool ok = p.face_off[0] == 0;
        for (int a = 0; a < p.n_poly && ok; a++) ok = p.face_off[a + 1] - p.face_off[a] >= 4;
        if (!ok) continue;  // build_model reports the invalid argument
        const int base = (int)(planes.size() / 4);
        for (int k = 0; k < p.face_off[p.n_poly]; k++) {  // normals normalised as se3gcopter_cpu.hpp:1116, then (n, n.p)
            const double* f = p.faces + 6 * (size_t)k;
            const double nn = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
            const double n0 = f[0] / nn, n1 = f[1] / nn, n2 = f[2] / nn;
            planes.push_back(n0), planes.push_back(n1), planes.push_back(n2);
            planes.push_back(n0 * f[3] + n1 * f[4] + n2 * f[5]);
        }
        for (int k = 0; k < 2 * p.n_poly - 1; k++) {
            const int a = k / 2;
            const int F = (k % 2 == 0) ? p.face_off[a + 1] - p.face_off[a] : p.face_off[a + 2] - p.face_off[a];  // [hPs[i], hPs[i+1]] contiguous
            polys.push_back(make_int2(base + p.face_off[a], F));
            vstart.push_back(slots);
            slots += 2 * F;
        }
    }
    first_poly[count] = (int)polys.size();
    const int npoly = (int)polys.size();
    if (npoly == 0) return FRX_OK;
    double *d_planes = nullptr, *d_verts = nullptr;
    int2* d_polys = nullptr;
    int *d_vstart = nullptr, *d_vcount = nullptr;
    std::vector<double> verts(3 * (size_t)slots);
    std::vector<int> vcount(npoly);
    bool ok = cudaMalloc(&d_planes, sizeof(double) * planes.size()) == cudaSuccess && cudaMalloc(&d_polys, sizeof(int2) * npoly) == cudaSuccess &&
              cudaMalloc(&d_vstart, sizeof(int) * npoly) == cudaSuccess && cudaMalloc(&d_vcount, sizeof(int) * npoly) == cudaSuccess &&
              cudaMalloc(&d_verts, sizeof(double) * verts.size()) == cudaSuccess;
    ok = ok && cudaMemcpy(d_planes, planes.data(), sizeof(double) * planes.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_polys, polys.data(), sizeof(int2) * npoly, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_vstart, vstart.data(), sizeof(int) * npoly, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && frx::launch_enumerate_vertices(d_planes, d_polys, d_vstart, npoly, d_verts, d_vcount, nullptr) == cudaSuccess;
    ok = ok && cudaMemcpy(vcount.data(), d_vcount, sizeof(int) * npoly, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(verts.data(), d_verts, sizeof(double) * verts.size(), cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d_planes), cudaFree(d_polys), cudaFree(d_vstart), cudaFree(d_vcount), cudaFree(d_verts);
    if (!ok) return cuda_fail(cudaGetLastError(), "device geometry");
    for (int i = 0; i < count; i++) {
        const int k0 = first_poly[i], k1 = first_poly[i + 1];
        if (k1 == k0) continue;
        bool fallback = false, infeasible = false;
        for (int k = k0; k < k1; k++) fallback |= vcount[k] == GEO_FALLBACK, infeasible |= vcount[k] == GEO_INFEASIBLE;
        if (fallback) continue;  // host enumeration inside build_model
        if (infeasible) {
            G.status[i] = FRX_ERR_INFEASIBLE_GEOMETRY;
            continue;
        }
        std::vector<int>& off = G.off[i];
        std::vector<double>& v = G.verts[i];
        off.assign(k1 - k0 + 1, 0);
        for (int k = k0; k < k1; k++) {
            const double* src = &verts[3 * (size_t)vstart[k]];
            const int V = vcount[k];
            for (int c = 0; c < V; c++)
                for (int d = 0; d < 3; d++) v.push_back(c == 0 ? src[d] : src[3 * c + d] - src[d]);  // origin + edges, :1049
            off[k - k0 + 1] = off[k - k0] + V;
        }
    }
    return FRX_OK;
}

extern "C" {

// ---- corridor generation -----------------------------------------------------------------------------
int frx_corridor_dilate(const double* obs, int n_obs, const double* p1, const double* p2, const double* bbox, double* faces, int cap_faces,
                        double* ellipsoid12) {
    if ((n_obs > 0 && !obs) || n_obs < 0 || !p1 || !p2 || !bbox || !faces) return FRX_ERR_INVALID_ARG;
    std::vector<double> f;
    const int F = dilate_segment(obs, n_obs, p1, p2, bbox, f, ellipsoid12);
    if (F > cap_faces) return FRX_ERR_CAPACITY;
    std::copy(f.begin(), f.end(), faces);
    return F;
}
int frx_corridor_generate(const double* obs, int n_obs, const double* path, int n_path, const double* bbox, double map_height,
                          frx_is_blocked_t blocked, void* ctx, int* face_off, double* faces, int cap_poly, int cap_faces, int* n_poly) {
    if ((n_obs > 0 && !obs) || n_obs < 0 || !path || n_path < 2 || !bbox || !face_off || !faces || !n_poly) return FRX_ERR_INVALID_ARG;
    std::vector<int> off;
    std::vector<double> f;
    const int P = generate_corridor(obs, n_obs, path, n_path, bbox, map_height, blocked, ctx, off, f);
    if (P > cap_poly || (int)f.size() / 6 > cap_faces) return FRX_ERR_CAPACITY;
    std::copy(off.begin(), off.end(), face_off);
    std::copy(f.begin(), f.end(), faces);
    *n_poly = P;
    return FRX_OK;
}
}  // extern "C"
struct frx_cloud {
    int device = 0, n = 0;
    double* obs = nullptr;
};
extern "C" {
int frx_cloud_create(const double* obs, int n_obs, int device_ordinal, frx_cloud** out) {
    if (!obs || n_obs < 1 || !out) return FRX_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "[frx] no CUDA device: this library has no CPU fallback\n");
        return FRX_ERR_CUDA;
    }
    frx_cloud* c = new (std::nothrow) frx_cloud();
    if (!c) return FRX_ERR_NOMEM;
    c->n = n_obs;
    if (device_ordinal < 0) {
        if (cudaGetDevice(&c->device) != cudaSuccess) {
            delete c;
            return FRX_ERR_CUDA;
        }
    } else
        c->device = device_ordinal;
    DeviceScope device_scope_(c->device);
    if (!device_scope_.ok || cudaMalloc(&c->obs, sizeof(double) * 3 * (size_t)n_obs) != cudaSuccess ||
        cudaMemcpy(c->obs, obs, sizeof(double) * 3 * (size_t)n_obs, cudaMemcpyHostToDevice) != cudaSuccess) {
        if (c->obs) cudaFree(c->obs);
        delete c;
        return cuda_fail(cudaGetLastError(), "frx_cloud_create");
    }
    *out = c;
    return FRX_OK;
}
void frx_cloud_destroy(frx_cloud* c) {
    if (!c) return;
    DeviceScope device_scope_(c->device);
    if (c->obs) cudaFree(c->obs);
    delete c;
}
int frx_cloud_dilate_batch(frx_cloud* c, const double* segments, int n_seg, const double* bbox, int* face_cnt, double* faces, double* ellipsoids) {
    if (!c || !segments || n_seg < 1 || !bbox || !face_cnt || !faces) return FRX_ERR_INVALID_ARG;
    ON_DEVICE(c->device);
    double *d_seg = nullptr, *d_faces = nullptr, *d_ell = nullptr;
    int *d_idx = nullptr, *d_cnt = nullptr;
    unsigned char *d_alive = nullptr, *d_in0 = nullptr;
    const size_t nf = (size_t)n_seg * FRX_CORRIDOR_MAX_FACES * 6;
    bool ok = cudaMalloc(&d_seg, sizeof(double) * 6 * (size_t)n_seg) == cudaSuccess && cudaMalloc(&d_faces, sizeof(double) * nf) == cudaSuccess &&
              cudaMalloc(&d_ell, sizeof(double) * 12 * (size_t)n_seg) == cudaSuccess && cudaMalloc(&d_cnt, sizeof(int) * n_seg) == cudaSuccess &&
              cudaMalloc(&d_idx, sizeof(int) * (size_t)n_seg * c->n) == cudaSuccess && cudaMalloc(&d_alive, (size_t)n_seg * c->n) == cudaSuccess &&
              cudaMalloc(&d_in0, (size_t)n_seg * c->n) == cudaSuccess;
    ok = ok && cudaMemcpy(d_seg, segments, sizeof(double) * 6 * (size_t)n_seg, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && frx::launch_dilate_batch(c->obs, c->n, d_seg, n_seg, bbox, d_idx, d_alive, d_in0, d_cnt, d_faces, d_ell, nullptr) == cudaSuccess;
    ok = ok && cudaMemcpy(face_cnt, d_cnt, sizeof(int) * n_seg, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(faces, d_faces, sizeof(double) * nf, cudaMemcpyDeviceToHost) == cudaSuccess;
    if (ok && ellipsoids) ok = cudaMemcpy(ellipsoids, d_ell, sizeof(double) * 12 * (size_t)n_seg, cudaMemcpyDeviceToHost) == cudaSuccess;
    const cudaError_t err = ok ? cudaSuccess : cudaGetLastError();
    cudaFree(d_seg), cudaFree(d_faces), cudaFree(d_ell), cudaFree(d_cnt), cudaFree(d_idx), cudaFree(d_alive), cudaFree(d_in0);
    if (!ok) return cuda_fail(err, "frx_cloud_dilate_batch");
    return FRX_OK;
}

int frx_cloud_generate_batch(frx_cloud* c, const double* paths, const int* path_off, int n_paths, const double* bbox, double map_height,
                             frx_is_blocked_t blocked, void* ctx, int* poly_off, int* face_off, double* faces, int cap_poly, int cap_faces) {
    if (!c || !paths || !path_off || n_paths < 1 || !bbox || !poly_off || !face_off || !faces) return FRX_ERR_INVALID_ARG;
    std::vector<int> cur(n_paths, 0), k_of(n_paths, 0);
    std::vector<char> walking(n_paths, 1);
    std::vector<std::vector<int>> offs(n_paths, std::vector<int>(1, 0));
    std::vector<std::vector<double>> fcs(n_paths);
    std::vector<double> seg, out((size_t)n_paths * FRX_CORRIDOR_MAX_FACES * 6);
    std::vector<int> who, cnt(n_paths);
    for (int p = 0; p < n_paths; p++) walking[p] = path_off[p + 1] - path_off[p] >= 2;
    for (;;) {
        seg.clear(), who.clear();
        for (int p = 0; p < n_paths; p++) {
            if (!walking[p]) continue;
            const double* path = paths + 3 * (size_t)path_off[p];
            const int np = path_off[p + 1] - path_off[p];
            k_of[p] = corridor_walk_target(path, np, cur[p], blocked, ctx, 4.0);
            seg.insert(seg.end(), path + 3 * (size_t)cur[p], path + 3 * (size_t)cur[p] + 3);
            seg.insert(seg.end(), path + 3 * (size_t)k_of[p], path + 3 * (size_t)k_of[p] + 3);
            who.push_back(p);
        }
        if (who.empty()) break;
        const int s = frx_cloud_dilate_batch(c, seg.data(), (int)who.size(), bbox, cnt.data(), out.data(), nullptr);
        if (s != FRX_OK) return s;
        for (size_t q = 0; q < who.size(); q++) {
            const int p = who[q];
            if (cnt[q] < 0) return FRX_ERR_CAPACITY;
            const double* f = &out[q * FRX_CORRIDOR_MAX_FACES * 6];
            const double* path = paths + 3 * (size_t)path_off[p];
            const int np = path_off[p + 1] - path_off[p];
            int next = corridor_walk_advance(path, np, cur[p], k_of[p], f, cnt[q]);
            fcs[p].insert(fcs[p].end(), f, f + 6 * (size_t)cnt[q]);
            const double ceil_floor[12] = {0, 0, 1, 0, 0, map_height, 0, 0, -1, 0, 0, 0};  // MinCoPlan_CPU.cpp:85-91
            fcs[p].insert(fcs[p].end(), ceil_floor, ceil_floor + 12);
            offs[p].push_back(offs[p].back() + cnt[q] + 2);
            if (next >= 0 && next <= cur[p]) next = cur[p] + 1;  // the reference's advance rule can stall (wp == i): move on
            if (next < 0 || next >= np - 1)
                walking[p] = 0;
            else
                cur[p] = next;
        }
    }
    int P = 0, F = 0;
    poly_off[0] = 0, face_off[0] = 0;
    for (int p = 0; p < n_paths; p++) {
        const int np = (int)offs[p].size() - 1;
        if (P + np > cap_poly || F + offs[p].back() > cap_faces) return FRX_ERR_CAPACITY;
        for (int k = 0; k < np; k++) face_off[P + k + 1] = F + offs[p][k + 1];
        std::copy(fcs[p].begin(), fcs[p].end(), faces + 6 * (size_t)F);
        P += np, F += offs[p].back();
        poly_off[p + 1] = P;
    }
    return FRX_OK;
}

int frx_device_numa_node(int device_ordinal) { return device_numa_node(device_ordinal); }

// Moves the CALLING thread (and the threads it creates later) to the CPUs of the GPU's NUMA node and makes that node the
// preferred one for its allocations: what a one-process-per-GPU launcher calls first.  Returns the node, or -1 (nothing changed).
int frx_bind_thread_to_device(int device_ordinal) {
    NumaScope s(device_ordinal, /*keep=*/true);
    return (s.moved || s.policy) ? s.node : -1;
}

const char* frx_version(void) { return "fast_racing_b200 0.1 (sm_100a)"; }

const char* frx_strerror(int s) {
    switch (s) {
    case FRX_OK: return "ok";
    case FRX_ERR_INVALID_ARG: return "invalid argument";
    case FRX_ERR_INFEASIBLE_GEOMETRY: return "infeasible geometry: a polytope or a consecutive intersection has no interior";
    case FRX_ERR_CUDA: return "CUDA error or no usable CUDA device (there is no CPU fallback)";
    case FRX_ERR_CAPACITY: return "problem exceeds a kernel capacity limit";
    case FRX_ERR_NOMEM: return "out of memory";
    default: return s <= -1001 && s >= -1024 ? "L-BFGS status (see lbfgs.hpp:148-205)" : "unknown";
    }
}

void frx_batch_destroy(frx_batch* b) {
    if (!b) return;
    DeviceScope device_scope_(b->device);
    if (b->solver) frx_fast::solver_destroy(b->solver);
    if (b->solver_x) frx_exact::solver_destroy(b->solver_x);
    for (auto& d : b->bufs) d.release();
    if (b->x_pin) cudaFreeHost(b->x_pin);
    if (b->g_pin) cudaFreeHost(b->g_pin);
    if (b->cost_pin) cudaFreeHost(b->cost_pin);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_k1) cudaEventDestroy(b->ev_k1);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->aux) cudaStreamDestroy(b->aux);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int frx_batch_create(const frx_problem* problems, int count, int device_ordinal, frx_batch** out) {
    if (!problems || count < 1 || !out) return FRX_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        std::fprintf(stderr, "[frx] no CUDA device: this library has no CPU fallback\n");
        return FRX_ERR_CUDA;
    }
    frx_batch* b = new (std::nothrow) frx_batch();
    if (!b) return FRX_ERR_NOMEM;
    if (device_ordinal < 0) {
        if (cudaGetDevice(&b->device) != cudaSuccess) {
            delete b;
            return FRX_ERR_CUDA;
        }
    } else
        b->device = device_ordinal;
    DeviceScope device_scope_(b->device);
    if (!device_scope_.ok) {
        delete b;
        return FRX_ERR_CUDA;
    }
    // ---- V-polytopes of every corridor: one launch on the device ----
    DeviceGeometry geom;
    {
        const int gs = device_geometry(problems, count, geom);
        if (gs != FRX_OK) {
            delete b;
            return gs;
        }
    }
    // ---- host models (index tables, gridMesh; geometry only for what the device declined), in parallel over problems ----
    b->models.resize(count);
    std::vector<int> status(count, FRX_OK);
    {
        const int nt = std::max(1, std::min<int>(std::thread::hardware_concurrency(), std::min(count, 32)));
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++)
            pool.emplace_back([&, t]() {
                for (int i = t; i < count; i += nt) {
                    if (geom.status[i] != FRX_OK) {
                        status[i] = geom.status[i];
                        continue;
                    }
                    frx_problem p = problems[i];
                    if (!geom.off[i].empty()) p.vert_off = geom.off[i].data(), p.verts = geom.verts[i].data();
                    status[i] = build_model(p, b->models[i]);
                }
            });
        for (auto& th : pool) th.join();
    }
    for (int i = 0; i < count; i++)
        if (status[i] != FRX_OK) {
            const int s = status[i];
            delete b;
            return s;
        }
    // ---- pack ----
    const int B = count;
    b->piece_off.assign(B + 1, 0);
    b->x_off.assign(B + 1, 0);
    b->coarse_off.assign(B + 1, 0);
    for (int i = 0; i < B; i++) {
        const TrajectoryModel& m = b->models[i];
        b->piece_off[i + 1] = b->piece_off[i] + m.fineN;
        b->x_off[i + 1] = b->x_off[i] + m.n();
        b->coarse_off[i + 1] = b->coarse_off[i] + m.coarseN;
        b->max_M = std::max(b->max_M, m.fineN);
        b->max_coarse = std::max(b->max_coarse, m.coarseN);
    }
    const int TP = b->piece_off[B];
    std::vector<int> dimT(B), intervals, piece_traj(TP), piece_poly(TP), cons(TP), pf_off(TP), pf_cnt(TP), wv_off(TP), wv_cnt(TP), wxi(TP);
    std::vector<DevParams> par(B);
    std::vector<double> headtail(18 * (size_t)B), faces, verts;
    int max_faces = 0, max_samples = 0;
    for (int i = 0; i < B; i++) {
        const TrajectoryModel& m = b->models[i];
        dimT[i] = m.dimT;
        DevParams& P = par[i];
        P.rho = m.rho, P.sumT = m.sumT;
        P.ell[0] = m.par.horiz_half_len, P.ell[1] = m.par.horiz_half_len, P.ell[2] = m.par.vert_half_len;
        P.margin = m.par.safe_margin;
        P.vmax_sqr = m.par.v_max * m.par.v_max;
        P.thrmin_sqr = m.par.thr_acc_min * m.par.thr_acc_min;
        P.thrmax_sqr = m.par.thr_acc_max * m.par.thr_acc_max;
        P.bdrmax_sqr = m.par.body_rate_max * m.par.body_rate_max;
        P.gacc = m.par.grav_acc;
        for (int q = 0; q < 4; q++) P.chi[q] = m.par.chi[q];
        P.softT = m.softT, P.c2 = m.c2;
        std::memcpy(&headtail[18 * (size_t)i], m.ini, 9 * sizeof(double));
        std::memcpy(&headtail[18 * (size_t)i + 9], m.fin, 9 * sizeof(double));
        intervals.insert(intervals.end(), m.intervals.begin(), m.intervals.end());
        const int face_base = (int)(faces.size() / 6);
        faces.insert(faces.end(), m.faces.begin(), m.faces.end());
        std::vector<int> vbase(m.vpolys.size());
        for (size_t v = 0; v < m.vpolys.size(); v++) {
            vbase[v] = (int)(verts.size() / 3);
            verts.insert(verts.end(), m.vpolys[v].v.begin(), m.vpolys[v].v.end());
        }
        for (int p = 0; p < m.fineN; p++) {
            const int gp = b->piece_off[i] + p;
            piece_traj[gp] = i;
            piece_poly[gp] = m.idxHs[p];
            cons[gp] = m.cons[p];
            pf_off[gp] = face_base + m.face_off[m.idxHs[p]];
            pf_cnt[gp] = m.face_off[m.idxHs[p] + 1] - m.face_off[m.idxHs[p]];
            max_faces = std::max(max_faces, pf_cnt[gp]);
            max_samples = std::max(max_samples, m.cons[p] + 1);
            if (p < m.fineN - 1) {
                wv_off[gp] = vbase[m.idxVs[p]];
                wv_cnt[gp] = m.vpolys[m.idxVs[p]].cols();
                wxi[gp] = m.xi_off[p];
            } else {
                wv_off[gp] = 0, wv_cnt[gp] = 0, wxi[gp] = m.dimP;
            }
        }
    }
    BatchView& bv = b->bv;
    bv.B = B, bv.total_pieces = TP, bv.total_x = b->x_off[B];
    bv.max_faces = max_faces, bv.max_samples = max_samples;
    bv.slot_list = nullptr, bv.n_slots = 0, bv.max_pieces = b->max_M;
    b->fwd_smem = frx_fast::forward_smem_bytes(b->max_M, b->max_coarse);
    b->bwd_smem = frx_fast::backward_smem_bytes(b->max_M, b->max_coarse);
    b->fwd_smem_x = frx_exact::forward_smem_bytes(b->max_M, b->max_coarse);
    b->bwd_smem_x = frx_exact::backward_smem_bytes(b->max_M, b->max_coarse);
    if (b->fwd_smem_x > 227 * 1024 || b->bwd_smem_x > 227 * 1024) {
        delete b;
        return FRX_ERR_CAPACITY;  // band system of one trajectory must fit the shared memory of one SM
    }
    auto fail = [&](cudaError_t e, const char* what) {
        int r = cuda_fail(e, what);
        frx_batch_destroy(b);
        return r;
    };
    cudaError_t e;
#define UP(vec, field)                                             \
    if ((e = b->upload(vec, &bv.field)) != cudaSuccess) return fail(e, "upload " #field)
    UP(b->piece_off, piece_off);
    UP(b->x_off, x_off);
    UP(b->coarse_off, coarse_off);
    UP(dimT, dimT);
    UP(par, par);
    UP(headtail, headtail);
    UP(intervals, intervals);
    UP(piece_traj, piece_traj);
    UP(piece_poly, piece_poly);
    UP(cons, cons);
    UP(pf_off, piece_face_off);
    UP(pf_cnt, piece_face_cnt);
    UP(wv_off, wp_vert_off);
    UP(wv_cnt, wp_vert_cnt);
    UP(wxi, wp_xi_off);
    UP(faces, faces);
    {   // single-precision copy for K2's prefilter: skip a face only when  n.pos + slack < d_lo  proves it inactive in fp64
        std::vector<float> ff(4 * (faces.size() / 6));
        for (size_t k = 0; k < faces.size() / 6; k++) {
            const double* f = &faces[6 * k];
            const double d = f[0] * f[3] + f[1] * f[4] + f[2] * f[5];
            ff[4 * k + 0] = (float)f[0], ff[4 * k + 1] = (float)f[1], ff[4 * k + 2] = (float)f[2];
            ff[4 * k + 3] = (float)(d - 1.0e-6 * std::fabs(d) - 1.0e-6);  // rounded down beyond any fp32 error of the test
        }
        UP(ff, faces_f32);
    }
    UP(verts, verts);
#undef UP
#define SC(n, field)                                                \
    if ((e = b->scratch((size_t)(n), &bv.field)) != cudaSuccess) return fail(e, "alloc " #field)
    SC(TP, T);
    SC(3 * (size_t)TP, q);
    SC(4 * (size_t)TP, wpw);
    SC(18 * (size_t)TP, c);
    SC(96 * (size_t)TP, lu);
    SC(18 * (size_t)TP, gdC);
    SC(TP, gdT);
    SC(TP, pen);
    SC(TP, energy);
#undef SC
    if ((e = b->scratch((size_t)bv.total_x, &b->x_dev)) != cudaSuccess) return fail(e, "alloc x");
    if ((e = b->scratch((size_t)bv.total_x, &b->g_dev)) != cudaSuccess) return fail(e, "alloc g");
    if ((e = b->scratch((size_t)B, &b->cost_dev)) != cudaSuccess) return fail(e, "alloc cost");
    if ((e = b->scratch((size_t)B, &b->final_cost_dev)) != cudaSuccess) return fail(e, "alloc final cost");
    {
        NumaScope numa(b->device);  // pinned host buffers on the GPU's own NUMA node (see NumaScope)
        if ((e = cudaMallocHost(&b->x_pin, sizeof(double) * bv.total_x)) != cudaSuccess) return fail(e, "pinned x");
        if ((e = cudaMallocHost(&b->g_pin, sizeof(double) * bv.total_x)) != cudaSuccess) return fail(e, "pinned g");
        if ((e = cudaMallocHost(&b->cost_pin, sizeof(double) * B)) != cudaSuccess) return fail(e, "pinned cost");
        // first touch inside the scope: the pages are placed where the policy / the running CPU say
        std::memset(b->x_pin, 0, sizeof(double) * bv.total_x);
        std::memset(b->g_pin, 0, sizeof(double) * bv.total_x);
        std::memset(b->cost_pin, 0, sizeof(double) * B);
    }
    if ((e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
    if ((e = cudaEventCreate(&b->ev0)) != cudaSuccess) return fail(e, "event");
    if ((e = cudaEventCreate(&b->ev1)) != cudaSuccess) return fail(e, "event");
    if ((e = cudaStreamCreateWithFlags(&b->aux, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "aux stream");
    if ((e = cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
    if ((e = cudaEventCreateWithFlags(&b->ev_k1, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
    if ((e = cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "event");
    if (const char* ns = getenv("FRX_EVAL_SUBBATCHES")) b->nsub = std::max(1, atoi(ns));
    if (const char* nc = getenv("FRX_E2E_CHUNKS")) b->e2e_chunks = std::max(1, atoi(nc));
    b->pl = frx_fast::plan_penalty(bv);
    if (b->pl.smem > 227 * 1024) {
        frx_batch_destroy(b);
        return FRX_ERR_CAPACITY;
    }
    if ((e = frx_fast::configure_kernels(b->fwd_smem, b->bwd_smem, b->pl.smem)) != cudaSuccess) return fail(e, "configure kernels");
    if ((e = frx_exact::configure_kernels(b->fwd_smem_x, b->bwd_smem_x, b->pl.smem)) != cudaSuccess) return fail(e, "configure kernels (reference order)");
    *out = b;
    return FRX_OK;
}

int frx_batch_count(const frx_batch* b) { return b ? b->bv.B : FRX_ERR_INVALID_ARG; }

int frx_batch_dims(const frx_batch* b, int traj, int* n, int* pieces, int* dim_t, int* dim_p) {
    if (!b || traj < 0 || traj >= b->bv.B) return FRX_ERR_INVALID_ARG;
    const TrajectoryModel& m = b->models[traj];
    if (n) *n = m.n();
    if (pieces) *pieces = m.fineN;
    if (dim_t) *dim_t = m.dimT;
    if (dim_p) *dim_p = m.dimP;
    return FRX_OK;
}
int frx_batch_x_offsets(const frx_batch* b, int* off) {
    if (!b || !off) return FRX_ERR_INVALID_ARG;
    std::copy(b->x_off.begin(), b->x_off.end(), off);
    return FRX_OK;
}
int frx_batch_piece_offsets(const frx_batch* b, int* off) {
    if (!b || !off) return FRX_ERR_INVALID_ARG;
    std::copy(b->piece_off.begin(), b->piece_off.end(), off);
    return FRX_OK;
}
int frx_batch_tables(const frx_batch* b, int traj, int* intervals, int* idx_hs, int* idx_vs, int* cons) {
    if (!b || traj < 0 || traj >= b->bv.B) return FRX_ERR_INVALID_ARG;
    const TrajectoryModel& m = b->models[traj];
    if (intervals) std::copy(m.intervals.begin(), m.intervals.end(), intervals);
    if (idx_hs) std::copy(m.idxHs.begin(), m.idxHs.end(), idx_hs);
    if (idx_vs) std::copy(m.idxVs.begin(), m.idxVs.end(), idx_vs);
    if (cons) std::copy(m.cons.begin(), m.cons.end(), cons);
    return FRX_OK;
}
int frx_batch_vpolys(const frx_batch* b, int traj, int* vert_off, double* verts) {
    if (!b || traj < 0 || traj >= b->bv.B || !vert_off) return FRX_ERR_INVALID_ARG;
    const TrajectoryModel& m = b->models[traj];
    int off = 0;
    for (size_t v = 0; v < m.vpolys.size(); v++) {
        vert_off[v] = off;
        if (verts) std::copy(m.vpolys[v].v.begin(), m.vpolys[v].v.end(), verts + 3 * (size_t)off);
        off += m.vpolys[v].cols();
    }
    vert_off[m.vpolys.size()] = off;
    return FRX_OK;
}

int frx_batch_initial_guess(frx_batch* b, double* x_host) {
    if (!b || !x_host) return FRX_ERR_INVALID_ARG;
    const int B = b->bv.B;
    const int nt = std::max(1, std::min<int>(std::thread::hardware_concurrency(), std::min(B, 32)));
    const char* env = getenv("FRX_HOST_INIT");
    if (env && env[0] == '1') {  // everything on the host (threaded over the batch)
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++)
            pool.emplace_back([&, t]() {
                for (int i = t; i < B; i += nt) initial_guess(b->models[i], x_host + b->x_off[i]);
            });
        for (auto& th : pool) th.join();
        return FRX_OK;
    }
    // setInitial + backwardT on the host (sequential in the polytopes, a few hundred flops per trajectory), then the M - 1
    // waypoint solves of backwardP per trajectory on the device, one thread each
    ON_DEVICE(b->device);
    const int TP = b->bv.total_pieces;
    std::vector<double> targets(3 * (size_t)TP, 0.0);
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; t++)
            pool.emplace_back([&, t]() {
                std::vector<double> innerP;
                for (int i = t; i < B; i += nt) {
                    initial_guess_prepare(b->models[i], x_host + b->x_off[i], innerP);
                    std::copy(innerP.begin(), innerP.end(), targets.begin() + 3 * (size_t)b->piece_off[i]);
                }
            });
        for (auto& th : pool) th.join();
    }
    std::vector<long long> work_off((size_t)TP);
    long long total = 0;
    for (int i = 0; i < B; i++) {
        const TrajectoryModel& m = b->models[i];
        for (int p = 0; p < m.fineN; p++) {
            work_off[(size_t)b->piece_off[i] + p] = total;
            if (p < m.fineN - 1) total += nls_work_doubles(m.vpolys[m.idxVs[p]].cols() - 1);
        }
    }
    double *d_targets = nullptr, *d_work = nullptr;
    long long* d_off = nullptr;
    std::vector<double> xd((size_t)b->bv.total_x);
    bool ok = cudaMalloc(&d_targets, sizeof(double) * targets.size()) == cudaSuccess && cudaMalloc(&d_off, sizeof(long long) * TP) == cudaSuccess &&
              cudaMalloc(&d_work, sizeof(double) * (size_t)std::max<long long>(total, 1)) == cudaSuccess;
    ok = ok && cudaMemcpy(d_targets, targets.data(), sizeof(double) * targets.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_off, work_off.data(), sizeof(long long) * TP, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && frx::launch_waypoint_backward(TP, b->bv.piece_traj, b->bv.x_off, b->bv.dimT, b->bv.wp_xi_off, b->bv.wp_vert_off, b->bv.wp_vert_cnt,
                                             b->bv.verts, d_targets, d_off, b->x_dev, d_work, b->stream) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(xd.data(), b->x_dev, sizeof(double) * xd.size(), cudaMemcpyDeviceToHost, b->stream) == cudaSuccess &&
         cudaStreamSynchronize(b->stream) == cudaSuccess;
    const cudaError_t err = ok ? cudaSuccess : cudaGetLastError();
    cudaFree(d_targets), cudaFree(d_off), cudaFree(d_work);
    if (!ok) return cuda_fail(err, "frx_batch_initial_guess");
    for (int i = 0; i < B; i++) {  // the xi blocks come from the device, the tau blocks were written above
        const TrajectoryModel& m = b->models[i];
        std::copy(xd.begin() + b->x_off[i] + m.dimT, xd.begin() + b->x_off[i + 1], x_host + b->x_off[i] + m.dimT);
    }
    return FRX_OK;
}

int frx_batch_eval_launches(const frx_batch* b) { return b ? 3 * ((b->nsub >= 2 && b->bv.B >= 2 * b->nsub) ? b->nsub : 1) : FRX_ERR_INVALID_ARG; }
void* frx_batch_stream(frx_batch* b) { return b ? (void*)b->stream : nullptr; }

int frx_batch_eval_device(frx_batch* b, const double* x_dev, double* cost_dev, double* grad_dev, void* stream) {
    if (!b || !x_dev || !cost_dev || !grad_dev) return FRX_ERR_INVALID_ARG;
    ON_DEVICE(b->device);
    cudaStream_t st = (cudaStream_t)stream;
    CK(b->eval_on(x_dev, cost_dev, grad_dev, st));
    return FRX_OK;
}

int frx_batch_eval(frx_batch* b, const double* x_host, double* cost_host, double* grad_host) {
    if (!b || !x_host || !cost_host || !grad_host) return FRX_ERR_INVALID_ARG;
    ON_DEVICE(b->device);
    const size_t nx = sizeof(double) * (size_t)b->bv.total_x;
    if (x_host != b->x_pin) std::memcpy(b->x_pin, x_host, nx);  // staging only for foreign (pageable) buffers
    CK(cudaEventRecord(b->ev0, b->stream));
    const int B = b->bv.B;
    const int nch = (B >= 64 && b->e2e_chunks > 1) ? std::min(b->e2e_chunks, B / 32) : 1;
    if (nch <= 1) {
        CK(cudaMemcpyAsync(b->x_dev, b->x_pin, nx, cudaMemcpyHostToDevice, b->stream));
        CK(b->eval_on(b->x_dev, b->cost_dev, b->g_dev, b->stream));
        CK(cudaMemcpyAsync(b->g_pin, b->g_dev, nx, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaMemcpyAsync(b->cost_pin, b->cost_dev, sizeof(double) * B, cudaMemcpyDeviceToHost, b->stream));
    } else {
        // chunked: H2D(c) -> kernels(c) -> D2H(c) alternate between two streams, so that the PCIe copies of one chunk
        // run under the kernels of the other
        CK(cudaEventRecord(b->ev_fork, b->stream));
        CK(cudaStreamWaitEvent(b->aux, b->ev_fork, 0));
        for (int c = 0; c < nch; c++) {
            const int t0 = (int)((long long)B * c / nch), t1 = (int)((long long)B * (c + 1) / nch);
            const size_t o0 = (size_t)b->x_off[t0], o1 = (size_t)b->x_off[t1];
            cudaStream_t s = (c & 1) ? b->aux : b->stream;
            CK(cudaMemcpyAsync(b->x_dev + o0, b->x_pin + o0, sizeof(double) * (o1 - o0), cudaMemcpyHostToDevice, s));
            CK(b->eval_window_on(b->x_dev, b->cost_dev, b->g_dev, s, t0, t1));
            CK(cudaMemcpyAsync(b->g_pin + o0, b->g_dev + o0, sizeof(double) * (o1 - o0), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(b->cost_pin + t0, b->cost_dev + t0, sizeof(double) * (t1 - t0), cudaMemcpyDeviceToHost, s));
        }
        CK(cudaEventRecord(b->ev_join, b->aux));
        CK(cudaStreamWaitEvent(b->stream, b->ev_join, 0));
    }
    CK(cudaEventRecord(b->ev1, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    b->last_ms = ms;
    if (grad_host != b->g_pin) std::memcpy(grad_host, b->g_pin, nx);
    if (cost_host != b->cost_pin) std::memcpy(cost_host, b->cost_pin, sizeof(double) * b->bv.B);
    return FRX_OK;
}

int frx_batch_host_buffers(frx_batch* b, double** x_pinned, double** cost_pinned, double** grad_pinned) {
    if (!b) return FRX_ERR_INVALID_ARG;
    if (x_pinned) *x_pinned = b->x_pin;
    if (cost_pinned) *cost_pinned = b->cost_pin;
    if (grad_pinned) *grad_pinned = b->g_pin;
    return FRX_OK;
}

double frx_eval(void* instance, const double* x, double* g, const int n) {
    frx_batch* b = (frx_batch*)instance;
    if (!b || b->bv.B != 1 || n != b->bv.total_x) return NAN;
    double cost = NAN;
    if (frx_batch_eval(b, x, &cost, g) != FRX_OK) return NAN;
    return cost;
}

double frx_batch_last_ms(const frx_batch* b) { return b ? b->last_ms : -1.0; }

// Average device time (ms) of each of the three evaluation kernels over `reps` evaluations at x_dev,
// CUDA events on the given stream.  ms_out[3] = {minco_forward, penalty, minco_backward}.
int frx_batch_profile_eval(frx_batch* b, const double* x_dev, double* cost_dev, double* grad_dev, void* stream, int reps,
                           double* ms_out) {
    if (!b || !x_dev || !cost_dev || !grad_dev || !ms_out || reps < 1) return FRX_ERR_INVALID_ARG;
    ON_DEVICE(b->device);
    if (b->exact)
        CK(frx_exact::profile_eval(b->bv, b->pl, b->fwd_smem_x, b->bwd_smem_x, x_dev, cost_dev, grad_dev, (cudaStream_t)stream, reps, ms_out));
    else
        CK(frx_fast::profile_eval(b->bv, b->pl, b->fwd_smem, b->bwd_smem, x_dev, cost_dev, grad_dev, (cudaStream_t)stream, reps, ms_out));
    return FRX_OK;
}
// Debug builds (-DFRX_PHASE_CLK): clock64() marks of one CTA of K1 / K3 (zeros otherwise).
int frx_debug_phase_clocks(long long* out96) {
    if (!out96) return FRX_ERR_INVALID_ARG;
    CK(frx_fast::fetch_phase_clocks(out96));
    return FRX_OK;
}
// Debug builds: per-CTA globaltimer spans of K1 / K3, out[2][3][4096] = {kernel}{start ns, end ns, SM id}{block}.
int frx_debug_cta_spans(unsigned long long* out) {
    if (!out) return FRX_ERR_INVALID_ARG;
    CK(frx_fast::fetch_cta_spans(out));
    return FRX_OK;
}
// Host-side view of the update kernel's shared-memory ring (no device needed): out[8] = {stages, doubles per stage, dynamic smem
// bytes, tile width, column tiles, rows per pass-2 chunk, streamed (0/1), fixed doubles} for a trajectory of n variables in a
// batch of `batch` trajectories whose largest has max_n.  tests/test_host_cpu.py checks the invariants the bulk copies rely on.
int frx_debug_update_ring(int max_n, int n, int batch, int* out8) {
    if (!out8 || max_n < 1 || n < 1 || n > max_n || batch < 1) return FRX_ERR_INVALID_ARG;
    frx_fast::debug_update_ring(max_n, n, batch, out8);
    return FRX_OK;
}
int frx_debug_update_clocks(long long* out16, int reset) {
    if (!out16) return FRX_ERR_INVALID_ARG;
    CK(frx_fast::fetch_update_clocks(out16, reset));
    return FRX_OK;
}
// Launch geometry of the penalty kernel: out[6] = {samples/thread, threads/piece, pieces/CTA, CTA threads, grid, smem bytes}
int frx_batch_penalty_launch(const frx_batch* b, long long* out) {
    if (!b || !out) return FRX_ERR_INVALID_ARG;
    out[0] = b->pl.spt, out[1] = b->pl.tpp, out[2] = b->pl.ppc, out[3] = b->pl.threads, out[4] = b->pl.grid, out[5] = (long long)b->pl.smem;
    return FRX_OK;
}
// Measured fp64 FMA peak (TFLOP/s) of the current device: the secondary roofline of this path.
double frx_measure_fp64_tflops(void) { return frx_fast::measure_fp64_tflops(); }

// Reference-order mode: evaluation and solver in the reference's operation order (no FMA contraction, sequential
// reductions, divisions as written): bit-identical to the CPU code, slow by construction (validation / reproducibility).
// Cross-rank arg-min for a C/C++ caller that runs one rank per GPU: the only collective of the path — one
// ncclAllGather of {cost, global index} (16 B per rank) on the batch's stream, then a local arg-min (ties -> lowest index).
// The NCCL library of the process is used (dlopen of libnccl.so.2; nothing is linked), `nccl_comm` is the caller's ncclComm_t.
int frx_batch_argmin_allgather(frx_batch* b, void* nccl_comm, int n_ranks, long long global_index_offset, long long* best_global_index,
                               double* best_cost) {
    if (!b || !nccl_comm || n_ranks < 1 || !b->have_final) return FRX_ERR_INVALID_ARG;
    typedef ncclResult_t (*allgather_t)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    static allgather_t allgather = nullptr;
    if (!allgather) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        allgather = h ? (allgather_t)dlsym(h, "ncclAllGather") : nullptr;
        if (!allgather) {
            std::fprintf(stderr, "[frx] frx_batch_argmin_allgather: no NCCL library in this process (%s)\n", dlerror());
            return FRX_ERR_CUDA;
        }
    }
    ON_DEVICE(b->device);
    int idx = -1;
    double val = 0.0;
    int s = frx_batch_argmin(b, &idx, &val);
    if (s != FRX_OK) return s;
    double* buf = nullptr;  // [2] mine | [2 * n_ranks] everybody's
    CK(cudaMalloc(&buf, sizeof(double) * 2 * (size_t)(n_ranks + 1)));
    const double mine[2] = {val, (double)(global_index_offset + idx)};  // indices below 2^53 are exact in a double
    std::vector<double> all(2 * (size_t)n_ranks);
    bool ok = cudaMemcpyAsync(buf, mine, sizeof(mine), cudaMemcpyHostToDevice, b->stream) == cudaSuccess;
    ok = ok && allgather(buf, buf + 2, 2, ncclDouble, (ncclComm_t)nccl_comm, b->stream) == ncclSuccess;
    ok = ok && cudaMemcpyAsync(all.data(), buf + 2, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, b->stream) == cudaSuccess &&
         cudaStreamSynchronize(b->stream) == cudaSuccess;
    cudaFree(buf);
    if (!ok) return cuda_fail(cudaGetLastError(), "frx_batch_argmin_allgather");
    int best = 0;
    for (int r = 1; r < n_ranks; r++)
        if (all[2 * r] < all[2 * best] || (all[2 * r] == all[2 * best] && all[2 * r + 1] < all[2 * best + 1])) best = r;
    if (best_cost) *best_cost = all[2 * best];
    if (best_global_index) *best_global_index = (long long)all[2 * best + 1];
    return FRX_OK;
}

int frx_batch_set_solver_slots(frx_batch* b, int slots) {
    if (!b || slots < 0) return FRX_ERR_INVALID_ARG;
    b->solver_slots = slots;
    return FRX_OK;
}

int frx_batch_set_reference_order(frx_batch* b, int on) {
    if (!b) return FRX_ERR_INVALID_ARG;
    b->exact = on != 0;
    return FRX_OK;
}

// Intermediate device quantities of the last evaluation (for component-level parity tests):
// T[pieces], q[3*pieces] (last slot unused), c[18*pieces], gdC_lambda[18*pieces], pen[pieces], energy[pieces], gdT[pieces]
int frx_batch_debug_fetch(frx_batch* b, double* T, double* q, double* c, double* lambda, double* pen, double* energy,
                          double* gdT) {
    if (!b) return FRX_ERR_INVALID_ARG;
    ON_DEVICE(b->device);
    CK(cudaStreamSynchronize(b->stream));
    const size_t TP = b->bv.total_pieces;
    if (T) CK(cudaMemcpy(T, b->bv.T, sizeof(double) * TP, cudaMemcpyDeviceToHost));
    if (q) CK(cudaMemcpy(q, b->bv.q, sizeof(double) * 3 * TP, cudaMemcpyDeviceToHost));
    if (c) CK(cudaMemcpy(c, b->bv.c, sizeof(double) * 18 * TP, cudaMemcpyDeviceToHost));
    if (lambda) CK(cudaMemcpy(lambda, b->bv.gdC, sizeof(double) * 18 * TP, cudaMemcpyDeviceToHost));
    if (pen) CK(cudaMemcpy(pen, b->bv.pen, sizeof(double) * TP, cudaMemcpyDeviceToHost));
    if (energy) CK(cudaMemcpy(energy, b->bv.energy, sizeof(double) * TP, cudaMemcpyDeviceToHost));
    if (gdT) CK(cudaMemcpy(gdT, b->bv.gdT, sizeof(double) * TP, cudaMemcpyDeviceToHost));
    return FRX_OK;
}

// Host-only geometry entry point (no GPU needed): vertices of one H-polytope, for the CPU test tier.
int frx_enumerate_vertices(const double* faces, int F, double* out, int cap) {
    if (!faces || F < 4) return FRX_ERR_INVALID_ARG;
    std::vector<double> f(faces, faces + 6 * (size_t)F), v;
    for (int k = 0; k < F; k++) {
        double* n = &f[6 * (size_t)k];
        const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        n[0] /= nn, n[1] /= nn, n[2] /= nn;
    }
    if (!enumerate_vertices(f.data(), F, v)) return FRX_ERR_INFEASIBLE_GEOMETRY;
    const int V = (int)v.size() / 3;
    for (int i = 0; i < V && i < cap; i++)
        for (int d = 0; d < 3; d++) out[3 * i + d] = v[3 * (size_t)i + d];
    return V;
}

// Host-only: index tables + starting point of one problem without touching the GPU (CPU test tier).
int frx_host_model(const frx_problem* p, int* n, int* pieces, int* dim_t, int* dim_p, int* intervals, int* idx_hs,
                   int* idx_vs, int* cons, double* x0) {
    if (!p) return FRX_ERR_INVALID_ARG;
    TrajectoryModel m;
    const int s = build_model(*p, m);
    if (s != FRX_OK) return s;
    if (n) *n = m.n();
    if (pieces) *pieces = m.fineN;
    if (dim_t) *dim_t = m.dimT;
    if (dim_p) *dim_p = m.dimP;
    if (intervals) std::copy(m.intervals.begin(), m.intervals.end(), intervals);
    if (idx_hs) std::copy(m.idxHs.begin(), m.idxHs.end(), idx_hs);
    if (idx_vs) std::copy(m.idxVs.begin(), m.idxVs.end(), idx_vs);
    if (cons) std::copy(m.cons.begin(), m.cons.end(), cons);
    if (x0) initial_guess(m, x0);
    return FRX_OK;
}
// Host-only: V-polytopes of one problem (sizes: vert_off[2*n_poly]; verts may be NULL).
int frx_host_vpolys(const frx_problem* p, int* vert_off, double* verts) {
    if (!p || !vert_off) return FRX_ERR_INVALID_ARG;
    TrajectoryModel m;
    const int s = build_model(*p, m);
    if (s != FRX_OK) return s;
    int off = 0;
    for (size_t v = 0; v < m.vpolys.size(); v++) {
        vert_off[v] = off;
        if (verts) std::copy(m.vpolys[v].v.begin(), m.vpolys[v].v.end(), verts + 3 * (size_t)off);
        off += m.vpolys[v].cols();
    }
    vert_off[m.vpolys.size()] = off;
    return FRX_OK;
}

int frx_batch_optimize(frx_batch* b, double rel_tol, double* durations, double* coeffs, double* jerk_cost,
                       double* final_cost, int* lbfgs_status, int* iterations, int* evaluations, double* x_final) {
    if (!b) return FRX_ERR_INVALID_ARG;
    ON_DEVICE(b->device);
    // starting point: setInitial + backwardT/P on the host, one H2D copy
    std::vector<double> x0((size_t)b->bv.total_x);
    int s = frx_batch_initial_guess(b, x0.data());
    if (s != FRX_OK) return s;
    if (b->exact && !b->solver_x) {
        s = frx_exact::solver_create(b->bv, b->x_off, &b->solver_x);
        if (s != FRX_OK) return s;
    }
    if (!b->exact && !b->solver) {
        s = frx_fast::solver_create(b->bv, b->x_off, &b->solver);
        if (s != FRX_OK) return s;
    }
    b->last_solver_exact = b->exact;
    SolverIO io;
    io.x0_host = x0.data();
    io.rel_tol = rel_tol;
    io.stream = b->stream;
    io.slots = b->solver_slots;
    if (const char* e = getenv("FRX_SOLVER_SLOTS")) io.slots = atoi(e);
    if (const char* e = getenv("FRX_SOLVER_MAX_TICKS")) io.max_ticks = atoi(e) > 0 ? atoi(e) : io.max_ticks;  // tests: truncated runs
    auto eval = [b](const double* x_dev, double* cost_dev, double* grad_dev, cudaStream_t st) {
        return b->eval_on(x_dev, cost_dev, grad_dev, st);
    };
    CK(cudaEventRecord(b->ev0, b->stream));
    s = b->exact ? frx_exact::solver_run(b->solver_x, b->bv, io, eval) : frx_fast::solver_run(b->solver, b->bv, io, eval);
    if (s != FRX_OK) return s;
    CK(cudaEventRecord(b->ev1, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    b->last_ms = ms;
    // outputs
    const int B = b->bv.B;
    const size_t TP = b->bv.total_pieces;
    s = b->exact ? frx_exact::solver_fetch(b->solver_x, final_cost, lbfgs_status, iterations, evaluations, x_final)
                 : frx_fast::solver_fetch(b->solver, final_cost, lbfgs_status, iterations, evaluations, x_final);
    if (s != FRX_OK) return s;
    CK(cudaMemcpy(b->final_cost_dev, b->exact ? frx_exact::solver_final_cost_dev(b->solver_x) : frx_fast::solver_final_cost_dev(b->solver),
                  sizeof(double) * B, cudaMemcpyDeviceToDevice));
    b->have_final = true;
    // final trajectory: the last forward pass of the solver was at the final iterate (solver_run
    // re-runs K1 on x_final, CPU.hpp:1258-1262), so T / c / energy are current.
    std::vector<double> T(TP), c(18 * TP), en(TP);
    CK(cudaMemcpy(T.data(), b->bv.T, sizeof(double) * TP, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c.data(), b->bv.c, sizeof(double) * 18 * TP, cudaMemcpyDeviceToHost));
    if (durations) std::copy(T.begin(), T.end(), durations);
    if (coeffs) {
        // Piece(T_i, b.block<6,3>(6i,0)^T.rowwise().reverse()) : 3x6 column-major, highest order first (CPU.hpp:561)
        for (size_t i = 0; i < TP; i++)
            for (int col = 0; col < 6; col++)
                for (int d = 0; d < 3; d++) coeffs[18 * i + 3 * col + d] = c[18 * i + 3 * (5 - col) + d];
    }
    if (jerk_cost) {
        // getTrajJerkCost (CPU.hpp:507-520) on the final coefficients, sequential sum as the reference
        for (int t = 0; t < B; t++) {
            double obj = 0.0;
            for (int i = b->piece_off[t]; i < b->piece_off[t + 1]; i++) {
                const double* cc = &c[18 * (size_t)i];
                const double T1 = T[i], T2 = T1 * T1, T3 = T2 * T1, T4 = T2 * T2, T5 = T4 * T1;
                const double q33 = cc[9] * cc[9] + cc[10] * cc[10] + cc[11] * cc[11];
                const double q44 = cc[12] * cc[12] + cc[13] * cc[13] + cc[14] * cc[14];
                const double q55 = cc[15] * cc[15] + cc[16] * cc[16] + cc[17] * cc[17];
                const double q43 = cc[12] * cc[9] + cc[13] * cc[10] + cc[14] * cc[11];
                const double q53 = cc[15] * cc[9] + cc[16] * cc[10] + cc[17] * cc[11];
                const double q54 = cc[15] * cc[12] + cc[16] * cc[13] + cc[17] * cc[14];
                obj += 36.0 * q33 * T1 + 144.0 * q43 * T2 + 192.0 * q44 * T3 + 240.0 * q53 * T3 + 720.0 * q54 * T4 + 720.0 * q55 * T5;
            }
            jerk_cost[t] = obj;
        }
    }
    return FRX_OK;
}

int frx_batch_solver_ticks(const frx_batch* b) {
    if (!b) return 0;
    if (b->last_solver_exact) return b->solver_x ? frx_exact::solver_ticks(b->solver_x) : 0;
    return b->solver ? frx_fast::solver_ticks(b->solver) : 0;
}

// Sampled lbfgs_update_kernel launches of the last frx_batch_optimize: out[3] = {device ms, algorithmic history bytes streamed, samples}
int frx_batch_solver_update_stats(const frx_batch* b, double* out) {
    if (!b || !out) return FRX_ERR_INVALID_ARG;
    int n = 0;
    out[0] = out[1] = out[2] = 0.0;
    if (b->last_solver_exact) {
        if (b->solver_x) frx_exact::solver_update_stats(b->solver_x, &out[0], &out[1], &n);
    } else if (b->solver)
        frx_fast::solver_update_stats(b->solver, &out[0], &out[1], &n);
    out[2] = n;
    return FRX_OK;
}

// FRX_LBFGS_CHECK=1 (set before the first optimize of the handle): largest relative deviation between the vector-free
// direction and the two-loop recursion over the last solve; -1 when the check was off.
int frx_batch_solver_direction_check(const frx_batch* b, double* max_rel) {
    if (!b || !max_rel) return FRX_ERR_INVALID_ARG;
    *max_rel = -1.0;
    if (!b->last_solver_exact && b->solver) return frx_fast::solver_direction_check(b->solver, max_rel);
    return FRX_OK;
}

int frx_batch_argmin(frx_batch* b, int* best_traj, double* best_cost) {
    if (!b || !b->have_final) return FRX_ERR_INVALID_ARG;
    ON_DEVICE(b->device);
    int idx = -1;
    double val = 0.0;
    int s = b->last_solver_exact ? frx_exact::solver_argmin(b->solver_x, b->final_cost_dev, b->bv.B, b->stream, &idx, &val)
                                 : frx_fast::solver_argmin(b->solver, b->final_cost_dev, b->bv.B, b->stream, &idx, &val);
    if (s != FRX_OK) return s;
    if (best_traj) *best_traj = idx;
    if (best_cost) *best_cost = val;
    return FRX_OK;
}

}  // extern "C"
