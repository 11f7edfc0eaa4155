// kanode_peer.cu — the data-parallel step's collective, fused into the packing kernel, over NVLink peer memory.
//
// One process per GPU: trajectories are sharded, parameters replicated, and the only exchange of a training
// step is the sum over ranks of [gradient sum (np) | loss sum | trajectory count].  For the small-model ensembles that is a
// few hundred doubles: the cost of a library all-reduce is its launch + rendezvous latency, paid once per 2 ms step.
// Here the exchange is part of the kernel that packs the sums:
//   * every rank owns a MAILBOX in its own HBM: pkt[2][world][CAP] 16-byte packets (2 = parity of the call count, so a rank
//     that is already in call e+1 never overwrites what a slower rank still reads of call e);
//   * the mailboxes are cudaMalloc'ed by the library and exported as CUDA IPC handles (kanode_peer_export); every rank maps
//     all of them (kanode_peer_attach) — on a B200 node these mappings are NVLink/NVSwitch peer memory;
//   * ONE kernel per step and rank (peer_pack_allreduce_kernel, one block, a thread per entry): the thread packs its entry as
//     {lo32, epoch, hi32, epoch} and stores that packet into slot `rank` of EVERY mailbox with one 16-byte store (remote
//     stores over NVLink).  Each 8-byte half carries its own copy of the epoch, so a packet validates itself: the receiver
//     spins on the packets of its own mailbox until both epochs match — no flag array, no memory fence anywhere (a system-scope
//     fence per step, or worse an acquire load in the spin loop, costs 0.2 - 1.2 ms at 8 ranks: measured) — and adds the
//     world's entries in rank order: every rank computes bit-identical sums.
// No host involvement, no second kernel, no NCCL.  A peer that never arrives poisons the result with NaN after ~20 s and
// sets an error word instead of hanging the GPU.
//
// Reference: none (the reference is single-process); the step being served is Zygote.gradient(loss, p) -> Flux.update!
// (Lotka-Volterra/LV_driver_KANODE.jl:284-287) under data parallelism.
#include <cstring>

#include "kanode_host.h"

namespace kanode {

constexpr int PEER_CAP = KANODE_PEER_MAX_ENTRIES;
constexpr int PEER_W = KANODE_PEER_MAX_WORLD;
constexpr size_t PEER_BOX_BYTES = sizeof(uint4) * 2 * PEER_W * PEER_CAP;

struct PeerBoxes {
    uint4* pkt[PEER_W];                    // pkt[r]: mailbox of rank r as mapped in this process
};

__device__ __forceinline__ void st_packet(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_packet(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// n = np + 2 entries <= PEER_CAP; one block; `ep` = low 32 bits of the call count, never 0.
template <class T>
__global__ void __launch_bounds__(256) peer_pack_allreduce_kernel(const T* __restrict__ g, const double* __restrict__ loss, double count,
                                                                  int np, PeerBoxes box, int rank, int world, unsigned ep, int par,
                                                                  double* __restrict__ out, int* __restrict__ err) {
    const int n = np + 2, tid = threadIdx.x;
    const size_t half = (size_t)par * PEER_W * PEER_CAP;
    // ---- pack + scatter: entry i of this rank goes to slot [par][rank][i] of every mailbox ----
    for (int i = tid; i < n; i += blockDim.x) {
        const double v = i < np ? (double)g[i] : (i == np ? *loss : count);
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
        const uint4 pk = make_uint4((unsigned)bits, ep, (unsigned)(bits >> 32), ep);
        for (int r = 0; r < world; ++r) st_packet(box.pkt[r] + half + (size_t)rank * PEER_CAP + i, pk);
    }
    // ---- gather: the world's packets of this epoch in the own mailbox, summed in rank order ----
    const uint4* mine = box.pkt[rank] + half;
    const unsigned long long t0 = globaltimer_ns();
    for (int i = tid; i < n; i += blockDim.x) {
        double acc = 0.0;
        bool ok = true;
        for (int r = 0; r < world && ok; ++r) {
            const uint4* p = mine + (size_t)r * PEER_CAP + i;
            uint4 pk = ld_packet(p);
            unsigned spins = 0;
            while (pk.y != ep || pk.w != ep) {
                if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 20000000000ull) { ok = false; break; }
                __nanosleep(100);
                pk = ld_packet(p);
            }
            acc += __longlong_as_double((long long)(((unsigned long long)pk.z << 32) | pk.x));
        }
        if (!ok) { *err = 1; acc = __longlong_as_double(0x7ff8000000000000ll); }
        out[i] = acc;
    }
}

void peer_release(kanode_handle* h) {
    for (int r = 0; r < PEER_W; ++r) {
        if (h->peer_map[r] && h->peer_map[r] != h->peer_box) cudaIpcCloseMemHandle(h->peer_map[r]);
        h->peer_map[r] = nullptr;
    }
    if (h->peer_box) cudaFree(h->peer_box);
    if (h->peer_err) cudaFree(h->peer_err);
    h->peer_box = nullptr; h->peer_err = nullptr; h->peer_rank = -1; h->peer_world = 0; h->peer_epoch = 0;
}

static int peer_enter(kanode_handle* h, const char* what) {
    if (!h) return fail(nullptr, KANODE_ERR_INVALID, "null handle");
    if (!h->children.empty()) return fail(h, KANODE_ERR_UNSUPPORTED, "%s: a multi-device handle sums its devices itself", what);
    CK(h, cudaSetDevice(h->device));
    return 0;
}

template <class T> int peer_pack_allreduce(kanode_handle* h, const T* d_grad, const double* d_loss, int64_t count, double* d_packed) {
    if (int rc = peer_enter(h, "kanode_pack_allreduce_dev")) return rc;
    if (!d_grad || !d_loss || !d_packed || count < 0) return fail(h, KANODE_ERR_INVALID, "bad arguments");
    if (h->peer_world < 1) return fail(h, KANODE_ERR_INVALID, "kanode_peer_attach first");
    if (h->np + 2 > (size_t)PEER_CAP) return fail(h, KANODE_ERR_UNSUPPORTED, "np + 2 = %zu entries exceed KANODE_PEER_MAX_ENTRIES", h->np + 2);
    PeerBoxes box{};
    for (int r = 0; r < h->peer_world; ++r) box.pkt[r] = reinterpret_cast<uint4*>(h->peer_map[r]);
    ++h->peer_epoch;                                                   // same sequence on every rank; parity = double buffer
    unsigned ep = (unsigned)(h->peer_epoch & 0xffffffffull);
    if (ep == 0) ep = 0x80000000u;                                     // 0 marks a never-written packet
    peer_pack_allreduce_kernel<T><<<1, 256, 0, h->stream>>>(d_grad, d_loss, (double)count, (int)h->np, box, h->peer_rank, h->peer_world,
                                                           ep, (int)(h->peer_epoch & 1ull), d_packed, h->peer_err);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

}  // namespace kanode

using namespace kanode;

extern "C" {

int kanode_peer_export(kanode_handle* h, void* ipc_handle) {
    if (int rc = peer_enter(h, "kanode_peer_export")) return rc;
    if (!ipc_handle) return fail(h, KANODE_ERR_INVALID, "null ipc_handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == KANODE_IPC_HANDLE_BYTES, "IPC handle size");
    CK(h, cudaStreamSynchronize(h->stream));
    peer_release(h);
    CK(h, cudaMalloc(&h->peer_box, PEER_BOX_BYTES));
    CK(h, cudaMemset(h->peer_box, 0, PEER_BOX_BYTES));
    CK(h, cudaMalloc(reinterpret_cast<void**>(&h->peer_err), sizeof(int)));
    CK(h, cudaMemset(h->peer_err, 0, sizeof(int)));
    CK(h, cudaDeviceSynchronize());                                    // the zeroed flags are in memory before any peer maps them
    cudaIpcMemHandle_t ipc;
    CK(h, cudaIpcGetMemHandle(&ipc, h->peer_box));
    std::memcpy(ipc_handle, &ipc, sizeof ipc);
    return 0;
}

int kanode_peer_attach(kanode_handle* h, int32_t rank, int32_t world, const void* ipc_handles) {
    if (int rc = peer_enter(h, "kanode_peer_attach")) return rc;
    if (world < 1 || world > PEER_W || rank < 0 || rank >= world) return fail(h, KANODE_ERR_INVALID, "rank %d / world %d out of range (<= %d)", rank, world, PEER_W);
    if (!h->peer_box) return fail(h, KANODE_ERR_INVALID, "kanode_peer_export first");
    if (world > 1 && !ipc_handles) return fail(h, KANODE_ERR_INVALID, "null ipc_handles");
    for (int r = 0; r < world; ++r) {
        if (r == rank) { h->peer_map[r] = h->peer_box; continue; }
        cudaIpcMemHandle_t ipc;
        std::memcpy(&ipc, static_cast<const char*>(ipc_handles) + (size_t)r * KANODE_IPC_HANDLE_BYTES, sizeof ipc);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(h, KANODE_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (peer access between the two GPUs is required)", r, cudaGetErrorString(e));
        }
        h->peer_map[r] = p;
    }
    h->peer_rank = rank; h->peer_world = world; h->peer_epoch = 0;
    return 0;
}

int kanode_pack_allreduce_dev(kanode_handle* h, const float* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed) {
    return peer_pack_allreduce<float>(h, d_grad_sum, d_loss_sum, count, d_packed);
}
int kanode_pack_allreduce_dev_f64(kanode_handle* h, const double* d_grad_sum, const double* d_loss_sum, int64_t count, double* d_packed) {
    return peer_pack_allreduce<double>(h, d_grad_sum, d_loss_sum, count, d_packed);
}

int kanode_peer_status(kanode_handle* h) {
    if (int rc = peer_enter(h, "kanode_peer_status")) return rc;
    if (!h->peer_err) return 0;
    CK(h, cudaStreamSynchronize(h->stream));
    int e = 0;
    CK(h, cudaMemcpy(&e, h->peer_err, sizeof e, cudaMemcpyDeviceToHost));
    if (e) return fail(h, KANODE_ERR_SOLVER, "peer all-reduce: a rank did not arrive within 20 s; the packed sums of that step are NaN");
    return 0;
}

}  // extern "C"
