// DEIM interpolation-index selection -- replaces the body of deim_interpolation_indices,
// /root/reference/src/deim.jl:9-28.
//
// Per greedy step l (reference lines in brackets) ONE kernel launch does everything:
//   streaming body   r_i = |u_l[i] - sum_j U[i,j] c_j| fused with the abs-argmax and the runner-up     [l.22-24]
//                    (deim_step_kernel; deim_step4_kernel forms the next 4 residual columns of a panel)
//   fused tail       executed by the LAST CTA of the grid to finish (atomic ticket):
//                      reduce the per-CTA candidates, gather the winner's rows                        [l.16-20]
//                      exchange the record with the other ranks: plain stores into every peer's mailbox over
//                      NVLink + a flag per (slot, rank), then poll the own flags (no NCCL call on the chain)
//                      pick the global winner (value, then LOWEST global index = Julia's first-index
//                      tie-break, NaN maximal), record the margin to the runner-up
//                      extend the LU factors of the panel's 16 x 16 and the sub-panel's 4 x 4 pivot blocks by
//                      one row (bordered update) and solve for the next step's coefficients            [l.21]
// The k x k system of deim.jl:21 is only needed once per 16-column panel: A = inv(P'U) is kept explicitly and
// extended per panel by the block-inverse formula with the panel's own factors (three small launches).
//
// The streaming path (MOR_B200_DEIM_BLOCK=0, k <= 16, or a basis the bulk copies cannot take) is the reference's
// own access pattern: one pass over columns 1..l per step, coefficients from the bordered LU of the k x k block.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "internal.h"

namespace mor {

struct Cand {
    double val;
    long long idx;
    double val2;   // largest value at any OTHER row seen by this candidate's owner (runner-up)
};

#define MOR_IDX_NONE LLONG_MAX

// true when candidate a beats candidate b (Julia argmax semantics + lowest-index tie-break)
__device__ __forceinline__ bool cand_better(double va, long long ia, double vb, long long ib) {
    if (ia == MOR_IDX_NONE) return false;
    if (ib == MOR_IDX_NONE) return true;
    const bool na = va != va, nb = vb != vb;
    if (na | nb) return (na & nb) ? (ia < ib) : na;
    return (va > vb) || (va == vb && ia < ib);
}

// merge candidate (ov, oi, os) into (v, i, s): the loser's best value competes for the runner-up slot
__device__ __forceinline__ void cand_merge(double& v, long long& i, double& s, double ov, long long oi, double os) {
    s = fmax(s, os);
    if (cand_better(ov, oi, v, i)) {
        if (i != MOR_IDX_NONE) s = fmax(s, v);
        v = ov;
        i = oi;
    } else if (oi != MOR_IDX_NONE) {
        s = fmax(s, ov);
    }
}

__device__ __forceinline__ void cand_warp_reduce(double& v, long long& i, double& s) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, off);
        const long long oi = __shfl_xor_sync(0xffffffffu, i, off);
        const double os = __shfl_xor_sync(0xffffffffu, s, off);
        cand_merge(v, i, s, ov, oi, os);
    }
}

constexpr int DEIM_THREADS = 256;
constexpr int DEIM_WARPS = DEIM_THREADS / 32;
constexpr int DEIM_BLOCK = 16;   // panel width of the blocked path (= the 16-column stages of deim_panel_kernel)
constexpr int DEIM_SUB = 4;      // sub-panel width inside a panel (deim_step4_kernel)
constexpr int DEIM_MAXK = 1024;
constexpr int DEIM_MAXRANKS = MBOX_MAXRANKS;
constexpr int DEIM_SLOTS = MBOX_SLOTS;    // mailbox ring: a rank is never more than one step ahead of its peers
constexpr int DEIM_REC_MAX = MBOX_REC_MAX;
// small (panel / sub-panel) LU state: Uf | L | Z (dim x dim, row-major with fixed stride) | c (dim)
constexpr int BST_DOUBLES = 3 * DEIM_BLOCK * DEIM_BLOCK + DEIM_BLOCK;
constexpr int SST_DOUBLES = 3 * DEIM_SUB * DEIM_SUB + DEIM_SUB;
constexpr int SBUF_DOUBLES = DEIM_REC_MAX + 5;   // >= record, >= BST + SST + 20

// block-wide reduce; result valid in thread 0
__device__ __forceinline__ void cand_block_reduce(double& v, long long& i, double& s, double* sv, long long* si,
                                                  double* ss) {
    cand_warp_reduce(v, i, s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sv[warp] = v;
        si[warp] = i;
        ss[warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
        v = lane < DEIM_WARPS ? sv[lane] : -1.0;
        i = lane < DEIM_WARPS ? si[lane] : MOR_IDX_NONE;
        s = lane < DEIM_WARPS ? ss[lane] : -1.0;
        cand_warp_reduce(v, i, s);
    }
}

__device__ __forceinline__ double2 ld_stream2(const double* p) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// running best / runner-up of one thread; rows arrive in ascending order so ties keep the earlier row
__device__ __forceinline__ void cand_take(double& bv, long long& bi, double& b2, double v, long long idx) {
    if (!(bv != bv)) {
        if ((v != v) || v > bv) {
            b2 = bv;
            bv = v;
            bi = idx;
        } else if (v > b2) {
            b2 = v;
        }
    }
}

// ---- peer exchange: self-validating 16-byte words (value split in two halves, each next to a 32-bit tag) ----------
// The same idea as NCCL's LL protocol: a word {lo, tag, hi, tag} needs no fence and no separate flag -- whichever
// 8-byte half arrives first over NVLink, the reader only accepts it once both tags carry this step's number.
__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"((unsigned)b), "r"(tag),
                 "r"((unsigned)(b >> 32)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4* p, unsigned tag, double& v, long long deadline) {
    unsigned a0, a1, a2, a3;
    for (;;) {
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "l"(p) : "memory");
        if (a1 == tag && a3 == tag) break;
        if (clock64() > deadline) return false;
    }
    v = __longlong_as_double((long long)(((unsigned long long)a2 << 32) | a0));
    return true;
}

// Everything the fused tail needs (passed by value to the step kernels).
struct TailArgs {
    unsigned* ticket;                 // last-CTA election
    Cand* cands;                      // one candidate per CTA of this launch
    // rows gathered at the local winner
    const double* U;                  // basis shard
    int64_t ldU;
    int k;                            // columns of the basis
    int krec;                         // the record carries U[idx, 0..krec): k on the streaming path, 0 on the blocked one
    const double* Pn;                 // residual panel (blocked path) or nullptr
    int64_t ldn;
    int bw;
    const double* P2;                 // sub-panel or nullptr
    int64_t ld2;
    int sw;
    int64_t row_offset;
    // exchange
    uint4* ll;                        // own mailbox: DEIM_SLOTS x DEIM_MAXRANKS x DEIM_REC_MAX words
    uint4* peer_ll[DEIM_MAXRANKS];    // the peers' mailboxes (NVLink peer memory)
    double* send;                     // NCCL fallback: the local record goes here instead
    int rank, nranks, rec, slot;
    unsigned tag;                     // this step's number (never 0)
    int mode;                         // 0: fused (local + peer exchange + global), 1: local part only (NCCL follows)
    // global part
    double* W;                        // streaming path: k x k rows of the basis at the chosen points (row-major)
    long long* idx_out;
    double* margins;
    double* values;                   // winning residual of every step
    int i;                            // greedy step
    int blocked;
    double *bst, *sst, *Dsub;         // panel / sub-panel LU state, next sub-panel's coefficients
    int j, jj, s0;                    // step inside the panel / sub-panel, first panel column of the sub-panel
    int last;                         // final greedy step of the call (a zero pivot no longer matters)
    int* status;
    unsigned long long* tail_ns;      // accumulated duration of the tails (globaltimer)
};

// One bordered (Doolittle) update of a small LU state by ONE warp: append row i (values w[0..dim)), then solve for
// the coefficients of column i + 1.  The greedy choice is partial pivoting, so no row exchanges are needed and the
// factorisation equals dgetrf's.  Z = inv(Uf) is kept explicitly: every phase is a set of independent dot products.
// S = row stride of the state.  Deliberately NOT inlined / unrolled and shared by the 16 x 16 and the 4 x 4 state:
// the tail runs once per launch in one CTA with a cold instruction cache, so what it costs is the number of
// instruction bytes fetched, not the number of instructions executed.
__device__ __noinline__ void bordered_update_warp(double* Uf, double* L, double* Z, double* c, int S, int i, int dim,
                                                  const double* w) {
    const int x = threadIdx.x & 31;
    if (x < i) {   // L row i:  ell[x] = sum_{t <= x} w[t] Z[t][x]
        double s = 0.0;
#pragma unroll 1
        for (int t = 0; t <= x; ++t) s = fma(w[t], Z[t * S + x], s);
        L[i * S + x] = s;
    }
    __syncwarp();
    if (x >= i && x < dim) {   // U-factor row i:  Uf[i][x] = w[x] - sum_{t < i} ell[t] Uf[t][x]
        double s = 0.0;
#pragma unroll 1
        for (int t = 0; t < i; ++t) s = fma(L[i * S + t], Uf[t * S + x], s);
        Uf[i * S + x] = w[x] - s;
    }
    __syncwarp();
    const double pivot = Uf[i * S + i];
    if (x < i) {   // Z column i:  Z[x][i] = -(sum_{t = x}^{i-1} Z[x][t] Uf[t][i]) / pivot
        double s = 0.0;
#pragma unroll 1
        for (int t = x; t < i; ++t) s = fma(Z[x * S + t], Uf[t * S + i], s);
        Z[x * S + i] = -s / pivot;
    } else if (x == i) {
        Z[i * S + i] = 1.0 / pivot;
    }
    __syncwarp();
    if (i + 1 < dim && x <= i) {   // c = Z[0:i+1, 0:i+1] * Uf[0:i+1, i+1]
        double s = 0.0;
#pragma unroll 1
        for (int t = x; t <= i; ++t) s = fma(Z[x * S + t], Uf[t * S + i + 1], s);
        c[x] = s;
    }
    __syncwarp();
}

// Shared scratch of the tail (one instance per step kernel).
struct TailSmem {
    double hv[DEIM_MAXRANKS], h2[DEIM_MAXRANKS];   // every rank's (value, runner-up)
    long long hi[DEIM_MAXRANKS];                   // ... and global index
    double st[BST_DOUBLES + SST_DOUBLES];          // the small LU states, prefetched while the candidates reduce
    long long widx;
    int win, fail;
};

// Pick the global winner from the headers (thread 0), record index / value / margin.
__device__ __forceinline__ void deim_pick_winner(const TailArgs& a, TailSmem& ts) {
    int win = 0;
    double bv = -1.0, b2 = -1.0;
    long long bi = MOR_IDX_NONE;
    for (int r = 0; r < a.nranks; ++r) {
        const long long before = bi;
        cand_merge(bv, bi, b2, ts.hv[r], ts.hi[r], ts.h2[r]);
        if (bi != before) win = r;
    }
    ts.win = win;
    a.idx_out[a.i] = bi + 1;   // 1-based, like Julia
    a.values[a.i] = bv;
    // relative gap to the runner-up: by how much this argmax is decided (0: exact tie / NaN / zero residual)
    a.margins[a.i] = (bv == bv && bv > 0.0 && b2 >= 0.0) ? (bv - b2) / bv : (bv > 0.0 && b2 < 0.0 ? 1.0 : 0.0);
}

// Last part of the tail: `wrec` (shared memory) is the winner's record, ts.st the small LU states.
__device__ void deim_tail_finish(const TailArgs& a, const double* wrec, TailSmem& ts) {
    const int tid = threadIdx.x, k = a.krec;
    for (int jx = tid; jx < k; jx += DEIM_THREADS) a.W[(size_t)a.i * k + jx] = wrec[2 + jx];
    if (!a.blocked) return;
    constexpr int BB = DEIM_BLOCK * DEIM_BLOCK;
    double* bs = ts.st;
    double* ss = ts.st + BST_DOUBLES;
    const int warp = tid >> 5;
    if (warp < 2) {   // warp 0: the panel's 16 x 16 state, warp 1: the sub-panel's 4 x 4 state -- one copy of the code
        const bool big = warp == 0;
        double* st = big ? bs : ss;
        const int S = big ? DEIM_BLOCK : DEIM_SUB, SQ = S * S;
        bordered_update_warp(st, st + SQ, st + 2 * SQ, st + 3 * SQ, S, big ? a.j : a.jj, big ? a.bw : a.sw,
                             wrec + 2 + k + (big ? 0 : DEIM_BLOCK));
    }
    __syncthreads();
    // the pivot is the residual at the chosen row: exactly zero <=> P'U singular (SingularException in the reference)
    if (tid == 0 && bs[a.j * DEIM_BLOCK + a.j] == 0.0 && !a.last && *a.status == 0) *a.status = a.i + 1;
    for (int e = tid; e < BST_DOUBLES; e += DEIM_THREADS) a.bst[e] = bs[e];
    if (tid < SST_DOUBLES) a.sst[tid] = ss[tid];
    // the next step opens a new sub-panel of this panel: its coefficients with respect to the s0n points chosen so
    // far in the panel, D (16 x 4 column-major, zero padded) = Zb[0:s0n, 0:s0n] * Ufb[0:s0n, s0n : s0n+swn]
    const int s0n = a.s0 + a.sw;
    if (a.jj + 1 == a.sw && s0n < a.bw && tid < DEIM_BLOCK * DEIM_SUB) {
        const int t = tid % DEIM_BLOCK, q = tid / DEIM_BLOCK;
        const int swn = a.bw - s0n < DEIM_SUB ? a.bw - s0n : DEIM_SUB;
        double s = 0.0;
        if (t < s0n && q < swn)
            for (int r = t; r < s0n; ++r) s = fma(bs[2 * BB + t * DEIM_BLOCK + r], bs[r * DEIM_BLOCK + s0n + q], s);
        a.Dsub[q * DEIM_BLOCK + t] = s;
    }
}

__device__ __forceinline__ void deim_prefetch_state(const TailArgs& a, TailSmem& ts) {
    if (!a.blocked) return;
    for (int e = threadIdx.x; e < BST_DOUBLES; e += DEIM_THREADS) ts.st[e] = __ldcg(&a.bst[e]);
    if (threadIdx.x < SST_DOUBLES) ts.st[BST_DOUBLES + threadIdx.x] = __ldcg(&a.sst[threadIdx.x]);
}

// The tail, entered by every thread of the LAST CTA of a step kernel.  sbuf: >= rec doubles of shared memory.
__device__ void deim_tail(const TailArgs& a, double* sbuf, double* sv, long long* si, double* ss2, TailSmem& ts) {
    const int tid = threadIdx.x, k = a.krec, rec = a.rec;
    unsigned long long t0 = 0;
    // every other CTA of this grid has exited: the next step's grid may take the idle SMs now (it waits for this
    // grid's completion before it reads anything)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        ts.fail = 0;
    }
    deim_prefetch_state(a, ts);
    // (a) this rank's winner over the per-CTA candidates of the launch
    // all loads first (independent L2 round trips), then the merges
    constexpr int PER = 5;   // covers grids up to 1280 CTAs in one round
    double cv[PER], c2[PER];
    long long ci[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int t = tid + u * DEIM_THREADS;
        const bool in = t < (int)gridDim.x;
        cv[u] = in ? __ldcg(&a.cands[t].val) : -1.0;
        ci[u] = in ? __ldcg(&a.cands[t].idx) : MOR_IDX_NONE;
        c2[u] = in ? __ldcg(&a.cands[t].val2) : -1.0;
    }
    double bv = -1.0, b2 = -1.0;
    long long bi = MOR_IDX_NONE;
#pragma unroll
    for (int u = 0; u < PER; ++u) cand_merge(bv, bi, b2, cv[u], ci[u], c2[u]);
    for (int t = tid + PER * DEIM_THREADS; t < (int)gridDim.x; t += DEIM_THREADS)
        cand_merge(bv, bi, b2, __ldcg(&a.cands[t].val), __ldcg(&a.cands[t].idx), __ldcg(&a.cands[t].val2));
    __syncthreads();   // sv/si/ss2 were used by the CTA's own reduction
    cand_block_reduce(bv, bi, b2, sv, si, ss2);
    if (tid == 0) {
        ts.widx = bi;
        sbuf[0] = bv;
        sbuf[1] = __longlong_as_double(bi);
        sbuf[rec - 1] = b2;
        ts.hv[a.rank] = bv;
        ts.hi[a.rank] = bi;
        ts.h2[a.rank] = b2;
    }
    __syncthreads();
    // (b) the record: [value, bits(global index), U[idx, 0..k), panel row (16), sub-panel row (4), runner-up]
    const long long g = ts.widx;
    for (int jx = tid; jx < k; jx += DEIM_THREADS)
        sbuf[2 + jx] = g == MOR_IDX_NONE ? 0.0 : __ldcg(&a.U[(g - a.row_offset) + (int64_t)jx * a.ldU]);
    if (a.blocked) {
        if (tid < DEIM_BLOCK)
            sbuf[2 + k + tid] = (g == MOR_IDX_NONE || tid >= a.bw || a.Pn == nullptr) ? 0.0 : __ldcg(&a.Pn[(g - a.row_offset) + (int64_t)tid * a.ldn]);
        if (tid >= 32 && tid < 32 + DEIM_SUB) {
            const int q = tid - 32;
            sbuf[2 + k + DEIM_BLOCK + q] = (g == MOR_IDX_NONE || q >= a.sw || a.P2 == nullptr) ? 0.0 : __ldcg(&a.P2[(g - a.row_offset) + (int64_t)q * a.ld2]);
        }
    }
    __syncthreads();
    if (a.mode == 1) {   // NCCL fallback: the host all-gathers `send` and launches deim_tail_global_kernel
        for (int e = tid; e < rec; e += DEIM_THREADS) a.send[e] = sbuf[e];
        if (tid == 0) {
            *a.ticket = 0;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            atomicAdd(a.tail_ns, t1 - t0);
        }
        return;
    }
    if (a.nranks > 1) {
        // (c) publish: this rank's record into every peer's mailbox, slot (step mod 4), region `rank`
        const size_t my_off = ((size_t)a.slot * DEIM_MAXRANKS + a.rank) * DEIM_REC_MAX;
        for (int e = tid; e < rec; e += DEIM_THREADS) {
            const double v = sbuf[e];
            for (int p = 0; p < a.nranks; ++p)
                if (p != a.rank) ll_store(a.peer_ll[p] + my_off + e, v, a.tag);
        }
        // headers of the other ranks (value, index, runner-up): one thread per rank
        const long long deadline = clock64() + (1LL << 32);   // ~2 s: a peer died; fail instead of hanging the GPU
        if (tid < a.nranks && tid != a.rank) {
            const uint4* src = a.ll + ((size_t)a.slot * DEIM_MAXRANKS + tid) * DEIM_REC_MAX;
            double v = -1.0, ib = 0.0, s2 = -1.0;
            const bool ok = ll_load(src, a.tag, v, deadline) && ll_load(src + 1, a.tag, ib, deadline) &&
                            ll_load(src + rec - 1, a.tag, s2, deadline);
            ts.hv[tid] = v;
            ts.hi[tid] = ok ? __double_as_longlong(ib) : MOR_IDX_NONE;
            ts.h2[tid] = s2;
            if (!ok) ts.fail = 1;
        }
        __syncthreads();
    }
    // (d) global winner; its record into sbuf
    if (tid == 0) deim_pick_winner(a, ts);
    __syncthreads();
    if (ts.win != a.rank) {
        const uint4* src = a.ll + ((size_t)a.slot * DEIM_MAXRANKS + ts.win) * DEIM_REC_MAX;
        const long long deadline = clock64() + (1LL << 32);
        for (int e = tid; e < rec; e += DEIM_THREADS) {
            double v = 0.0;
            if (!ll_load(src + e, a.tag, v, deadline)) ts.fail = 1;
            sbuf[e] = v;
        }
    }
    __syncthreads();
    if (tid == 0 && ts.fail) atomicExch(a.status, -1);
    // (e) factor updates, next coefficients
    deim_tail_finish(a, sbuf, ts);
    if (tid == 0) {
        *a.ticket = 0;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        atomicAdd(a.tail_ns, t1 - t0);
    }
}

// elect the last CTA of the grid to finish its streaming work
__device__ __forceinline__ bool deim_last_cta(unsigned* ticket) {
    __shared__ unsigned last_s;
    __threadfence();
    if (threadIdx.x == 0) last_s = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (last_s) __threadfence();
    return last_s != 0;
}

// NCCL fallback: the global part of the tail as its own launch; `gathered` = nranks plain records.
__global__ void __launch_bounds__(DEIM_THREADS) deim_tail_global_kernel(TailArgs a, const double* gathered) {
    __shared__ double sbuf[SBUF_DOUBLES];
    __shared__ TailSmem ts;
    const int tid = threadIdx.x, rec = a.rec;
    deim_prefetch_state(a, ts);
    if (tid < a.nranks) {
        ts.hv[tid] = gathered[(size_t)tid * rec];
        ts.hi[tid] = __double_as_longlong(gathered[(size_t)tid * rec + 1]);
        ts.h2[tid] = gathered[(size_t)tid * rec + rec - 1];
    }
    __syncthreads();
    if (tid == 0) deim_pick_winner(a, ts);
    __syncthreads();
    for (int e = tid; e < rec; e += DEIM_THREADS) sbuf[e] = gathered[(size_t)ts.win * rec + e];
    __syncthreads();
    deim_tail_finish(a, sbuf, ts);
}

// Fused residual + abs-argmax over this rank's rows, then the tail.  l = number of previous columns.
// VEC: 16-byte loads (needs 16B-aligned base and even ld).  Each CTA walks chunks of
// 2*R*256 consecutive rows; for every column the CTA reads one contiguous 4*R KB run.
template <int R, bool VEC>
__global__ void __launch_bounds__(DEIM_THREADS)
deim_step_kernel(const double* __restrict__ U, int64_t m, int64_t ld, int l, const double* __restrict__ cg,
                 int64_t row_offset, TailArgs ta) {
    __shared__ double sc[SBUF_DOUBLES];   // coefficients during the streaming loop, record / LU state in the tail
    __shared__ double sv[DEIM_WARPS], ss2[DEIM_WARPS];
    __shared__ long long si[DEIM_WARPS];
    const int tid = threadIdx.x;
    // Programmatic dependent launch: this grid may have been scheduled while the previous step's tail was still
    // running in its last CTA; nothing of that step (coefficients, ticket) is touched before it has completed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int j = tid; j < l; j += DEIM_THREADS) sc[j] = cg[j];
    __syncthreads();

    const int64_t CH = 2 * R * DEIM_THREADS;
    const int64_t nchunks = (m + CH - 1) / CH;
    const double* __restrict__ ul = U + (int64_t)l * ld;
    double bv = -1.0, b2 = -1.0;
    long long bi = MOR_IDX_NONE;

    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t c0 = chunk * CH;
        if (c0 + CH <= m) {
            if (VEC) {
                double2 acc[R];
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = make_double2(0.0, 0.0);
                const double* p = U + c0 + 2 * tid;
                int j = 0;
                // two columns per trip: 2*R independent 16-byte loads in flight per thread
                for (; j + 2 <= l; j += 2) {
                    double2 v0[R], v1[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) v0[r] = ld_stream2(p + r * (2 * DEIM_THREADS));
#pragma unroll
                    for (int r = 0; r < R; ++r) v1[r] = ld_stream2(p + ld + r * (2 * DEIM_THREADS));
                    const double ca = sc[j], cb = sc[j + 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r].x = fma(v0[r].x, ca, acc[r].x);
                        acc[r].y = fma(v0[r].y, ca, acc[r].y);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r].x = fma(v1[r].x, cb, acc[r].x);
                        acc[r].y = fma(v1[r].y, cb, acc[r].y);
                    }
                    p += 2 * ld;
                }
                if (j < l) {
                    const double cj = sc[j];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double2 v = ld_stream2(p + r * (2 * DEIM_THREADS));
                        acc[r].x = fma(v.x, cj, acc[r].x);
                        acc[r].y = fma(v.y, cj, acc[r].y);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int64_t row = c0 + 2 * tid + r * (2 * DEIM_THREADS);
                    const double2 u = ld_stream2(ul + row);
                    cand_take(bv, bi, b2, fabs(u.x - acc[r].x), row_offset + row);
                    cand_take(bv, bi, b2, fabs(u.y - acc[r].y), row_offset + row + 1);
                }
            } else {
                double acc[2 * R];
#pragma unroll
                for (int r = 0; r < 2 * R; ++r) acc[r] = 0.0;
                const double* p = U + c0 + tid;
#pragma unroll 2
                for (int j = 0; j < l; ++j) {
                    const double cj = sc[j];
#pragma unroll
                    for (int r = 0; r < 2 * R; ++r)
                        acc[r] = fma(ld_stream1(p + r * DEIM_THREADS), cj, acc[r]);
                    p += ld;
                }
#pragma unroll
                for (int r = 0; r < 2 * R; ++r) {
                    const int64_t row = c0 + tid + r * DEIM_THREADS;
                    cand_take(bv, bi, b2, fabs(ld_stream1(ul + row) - acc[r]), row_offset + row);
                }
            }
        } else {  // ragged last chunk
            for (int64_t row = c0 + tid; row < m; row += DEIM_THREADS) {
                double acc = 0.0;
                const double* p = U + row;
                for (int j = 0; j < l; ++j) acc = fma(p[(int64_t)j * ld], sc[j], acc);
                cand_take(bv, bi, b2, fabs(ul[row] - acc), row_offset + row);
            }
        }
    }
    cand_block_reduce(bv, bi, b2, sv, si, ss2);
    if (tid == 0) {
        ta.cands[blockIdx.x].val = bv;
        ta.cands[blockIdx.x].idx = bi;
        ta.cands[blockIdx.x].val2 = b2;
    }
    if (!deim_last_cta(ta.ticket)) return;
    __shared__ TailSmem ts;
    deim_tail(ta, sc, sv, si, ss2, ts);
}

// Second blocking level: residuals of the next sw <= 4 columns of a panel P1 with respect to the s0 <= 12
// interpolation points already chosen inside that panel,
//     R2[:, q] = P1[:, s0 + q] - sum_{t < s0} P1[:, t] * D[t][q],
// written to the sub-panel R2 and fused with the abs-argmax of its first column (= the next greedy step) and the
// tail.  One streaming pass over s0 + sw panel columns instead of one pass per greedy step.  D: 16 x 4 column-major.
template <int R>
__global__ void __launch_bounds__(DEIM_THREADS)
deim_step4_kernel(const double* __restrict__ P1, int64_t m, int64_t ld1, int s0, int sw, const double* __restrict__ Dg,
                  double* __restrict__ R2, int64_t ld2, int64_t row_offset, TailArgs ta) {
    __shared__ double sbuf[SBUF_DOUBLES];
    __shared__ double sv[DEIM_WARPS], ss2[DEIM_WARPS];
    __shared__ long long si[DEIM_WARPS];
    double* sd = sbuf;
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // see deim_step_kernel
    if (tid < DEIM_BLOCK * DEIM_SUB) sd[tid] = Dg[tid];
    __syncthreads();

    const int64_t CH = 2 * R * DEIM_THREADS;
    const int64_t nchunks = (m + CH - 1) / CH;
    double bv = -1.0, b2 = -1.0;
    long long bi = MOR_IDX_NONE;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        const int64_t c0 = chunk * CH;
        if (c0 + CH <= m) {
            double2 acc[R][DEIM_SUB];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int q = 0; q < DEIM_SUB; ++q) acc[r][q] = make_double2(0.0, 0.0);
            const double* p = P1 + c0 + 2 * tid;
#pragma unroll 2
            for (int t = 0; t < s0; ++t) {
                double2 v[R];
#pragma unroll
                for (int r = 0; r < R; ++r) v[r] = ld_stream2(p + r * (2 * DEIM_THREADS));
#pragma unroll
                for (int q = 0; q < DEIM_SUB; ++q) {
                    const double d = sd[q * DEIM_BLOCK + t];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        acc[r][q].x = fma(v[r].x, d, acc[r][q].x);
                        acc[r][q].y = fma(v[r].y, d, acc[r][q].y);
                    }
                }
                p += ld1;
            }
#pragma unroll
            for (int q = 0; q < DEIM_SUB; ++q) {
                if (q < sw) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int64_t row = c0 + 2 * tid + r * (2 * DEIM_THREADS);
                        const double2 u = ld_stream2(p + r * (2 * DEIM_THREADS));
                        const double2 res = make_double2(u.x - acc[r][q].x, u.y - acc[r][q].y);
                        *reinterpret_cast<double2*>(R2 + (int64_t)q * ld2 + row) = res;
                        if (q == 0) {
                            cand_take(bv, bi, b2, fabs(res.x), row_offset + row);
                            cand_take(bv, bi, b2, fabs(res.y), row_offset + row + 1);
                        }
                    }
                }
                p += ld1;
            }
        } else {  // ragged last chunk
            for (int64_t row = c0 + tid; row < m; row += DEIM_THREADS) {
                double acc[DEIM_SUB] = {0.0, 0.0, 0.0, 0.0};
                const double* p = P1 + row;
                for (int t = 0; t < s0; ++t) {
                    const double v = p[(int64_t)t * ld1];
#pragma unroll
                    for (int q = 0; q < DEIM_SUB; ++q) acc[q] = fma(v, sd[q * DEIM_BLOCK + t], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < DEIM_SUB; ++q) {
                    if (q < sw) {
                        const double res = p[(int64_t)(s0 + q) * ld1] - acc[q];
                        R2[(int64_t)q * ld2 + row] = res;
                        if (q == 0) cand_take(bv, bi, b2, fabs(res), row_offset + row);
                    }
                }
            }
        }
    }
    cand_block_reduce(bv, bi, b2, sv, si, ss2);
    if (tid == 0) {
        ta.cands[blockIdx.x].val = bv;
        ta.cands[blockIdx.x].idx = bi;
        ta.cands[blockIdx.x].val2 = b2;
    }
    if (!deim_last_cta(ta.ticket)) return;   // includes the fence that publishes this CTA's R2 stores
    __shared__ TailSmem ts;
    deim_tail(ta, sbuf, sv, si, ss2, ts);
}

__device__ __forceinline__ double warp_sum(double s) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
}

// ---- streaming path: bordered LU of the k x k block P'U, one row per greedy step -----------------------------
// Replicated on every rank: append row i of W = P'U (already stored by the tail), extend the LU factors (L row i,
// U-factor row i, Z column i) and compute the coefficient vector c (length i+1) of the NEXT step.  All k x k arrays
// are row-major with stride k; Zt is the transpose of Z, kept so that every phase reads with the output index
// contiguous.  Every phase is a (triangular) matrix-vector product: thread (x, y) accumulates the terms
// t = y (mod ways) of output x, the `ways` partial sums are combined through shared memory -- a sequential depth of
// k / ways instead of k on this latency-critical single-CTA kernel.
__global__ void __launch_bounds__(1024)
deim_update_kernel(const double* __restrict__ Wm, double* __restrict__ Uf, double* __restrict__ Lm, double* __restrict__ Z,
                   double* __restrict__ Zt, double* __restrict__ c, int i, int k, int last, int* __restrict__ status) {
    __shared__ double w[1024], ell[1024], ucol[1024], y[1024], part[1024];
    const int tid = threadIdx.x;
    const int KP = (k + 31) & ~31;       // outputs per phase, padded to whole warps
    const int ways = 1024 / KP;          // reduction split (k <= 1024 => ways >= 1)
    const int x = tid % KP, yy = tid / KP;
    const bool active = yy < ways;
    for (int j = tid; j < k; j += 1024) w[j] = Wm[(size_t)i * k + j];
    __syncthreads();
    // combine the `ways` partial sums of output x (valid in threads with yy == 0)
    auto combine = [&](double s) -> double {
        part[tid] = s;
        __syncthreads();
        double tot = 0.0;
        if (yy == 0)
            for (int q = 0; q < ways; ++q) tot += part[q * KP + x];
        __syncthreads();
        return tot;
    };
    // L row: ell[j] = sum_{t <= j} w[t] * Z[t][j],  j < i   (Z = inv(Uf[0:i,0:i]), upper triangular)
    {
        double s = 0.0;
        if (active && x < i) {
#pragma unroll 8
            for (int t = yy; t <= x; t += ways) s = fma(w[t], __ldcg(&Z[(size_t)t * k + x]), s);
        }
        s = combine(s);
        if (yy == 0 && x < i) {
            ell[x] = s;
            Lm[(size_t)i * k + x] = s;
        }
    }
    __syncthreads();
    // U-factor row i:  Uf[i][j] = w[j] - sum_{t<i} ell[t] * Uf[t][j],  j >= i
    {
        const int j = i + x;
        double s = 0.0;
        if (active && j < k) {
#pragma unroll 8
            for (int t = yy; t < i; t += ways) s = fma(ell[t], __ldcg(&Uf[(size_t)t * k + j]), s);
        }
        s = combine(s);
        if (yy == 0 && j < k) Uf[(size_t)i * k + j] = w[j] - s;
    }
    __syncthreads();
    for (int t = tid; t < i; t += 1024) ucol[t] = Uf[(size_t)t * k + i];
    __syncthreads();
    const double pivot = Uf[(size_t)i * k + i];
    if (tid == 0 && pivot == 0.0 && !last && *status == 0) *status = i + 1;
    // Z column i:  Z[r][i] = -(sum_{t=r}^{i-1} Z[r][t] * Uf[t][i]) / pivot,  r < i;  Z[i][i] = 1/pivot
    {
        double s = 0.0;
        if (active && x < i) {
#pragma unroll 8
            for (int t = x + yy; t < i; t += ways) s = fma(__ldcg(&Zt[(size_t)t * k + x]), ucol[t], s);
        }
        s = combine(s);
        if (yy == 0 && x <= i) {
            const double z = (x == i) ? 1.0 / pivot : -s / pivot;
            Z[(size_t)x * k + i] = z;
            Zt[(size_t)i * k + x] = z;
        }
    }
    if (i + 1 >= k) return;
    __syncthreads();
    for (int t = tid; t <= i; t += 1024) y[t] = Uf[(size_t)t * k + i + 1];
    __syncthreads();
    // c = Z[0:i+1, 0:i+1] * y   (y = inv(L) * P'u_{l+1} is column i+1 of the U-factor)
    {
        double s = 0.0;
        if (active && x <= i) {
#pragma unroll 8
            for (int t = x + yy; t <= i; t += ways) s = fma(Zt[(size_t)t * k + x], y[t], s);   // row i was written above
        }
        s = combine(s);
        if (yy == 0 && x <= i) c[x] = s;
    }
}

// ---- blocked path: A = inv(P'U) kept explicitly, extended once per panel ----------------------------------------
// The k x k block P'U is never formed on this path.  At a panel boundary two pieces of it are fetched from the
// basis (each rank contributes the rows it owns, zeros elsewhere; the host all-reduces them when there are several):
//   W21[c][t] = U[idx[l0 + c], t],       t < l0              (rows of the panel's new points, earlier columns)
//   W12[t][q] = U[idx[t], l1 + q],       t < l1 = l0 + bw    (rows of all points so far, the next panel's columns)
// One launch writes both (zeros for rows owned by other ranks and for the padding columns of W12):
//   W21 (bw x l0, row-major, stride l0) then W12 (l1 x 16, row-major)
__global__ void __launch_bounds__(256)
deim_gather_kernel(const double* __restrict__ U, int64_t m, int64_t ld, int64_t row_offset, const long long* __restrict__ idx,
                   int l0, int bw, int bwn, double* __restrict__ W21, double* __restrict__ W12) {
    const int l1 = l0 + bw, n21 = bw * l0;
    int e = blockIdx.x * 256 + threadIdx.x;
    if (e < n21) {
        const int c = e / l0, t = e % l0;
        const long long loc = idx[l0 + c] - 1 - row_offset;
        W21[e] = (loc >= 0 && loc < m) ? U[loc + (int64_t)t * ld] : 0.0;
        return;
    }
    e -= n21;
    if (e >= l1 * DEIM_BLOCK) return;
    const int t = e / DEIM_BLOCK, q = e % DEIM_BLOCK;
    const long long loc = idx[t] - 1 - row_offset;
    W12[e] = (q < bwn && loc >= 0 && loc < m) ? U[loc + (int64_t)(l1 + q) * ld] : 0.0;
}

// Coefficients of the panel that starts at column l0:  C = A[0:l0, 0:l0] * W12  (column q of C solves
// (P'U_l0) c = P'u_{l0+q}, deim.jl:21).  W12[s][q] = Wsrc[s * ldw + q].  One warp per row t of C: the lanes split the
// dot products (row t of A is contiguous, row s of W12 is 16 contiguous doubles), 16 partial sums per lane are
// combined by shuffles -- a latency chain of l0 / 32 loads instead of l0.  Written twice: packed for
// deim_panel_kernel, Cp[c][q][kk] = C[16 c + kk][q] (the host zeroes the padding once), and plain (C12[t * 16 + q])
// for the panel-end update.
__global__ void __launch_bounds__(256)
deim_coef_kernel(const double* __restrict__ A, int k, int l0, int bw, const double* __restrict__ Wsrc, int ldw,
                 double* __restrict__ Cp, double* __restrict__ C12) {
    const int lane = threadIdx.x & 31, t = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= l0) return;
    double acc[DEIM_BLOCK];
#pragma unroll
    for (int q = 0; q < DEIM_BLOCK; ++q) acc[q] = 0.0;
    const double* a = A + (size_t)t * k;
    for (int r = lane; r < l0; r += 32) {
        const double av = __ldcg(&a[r]);
        const double* w = Wsrc + (size_t)r * ldw;
#pragma unroll
        for (int q = 0; q < DEIM_BLOCK; ++q) acc[q] = fma(av, q < bw ? __ldcg(&w[q]) : 0.0, acc[q]);
    }
#pragma unroll
    for (int q = 0; q < DEIM_BLOCK; ++q) acc[q] = warp_sum(acc[q]);
    if (lane < DEIM_BLOCK) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < DEIM_BLOCK; ++q) v = (q == lane) ? acc[q] : v;
        Cp[(size_t)(t / DEIM_BLOCK) * (DEIM_BLOCK * 20) + lane * 20 + (t % DEIM_BLOCK)] = v;
        C12[t * DEIM_BLOCK + lane] = v;
    }
}

// Panel end, steps 1 + 2.  Every CTA first forms Sinv = inv(S) = Z * inv(L) of the panel's pivot block S = P'R = L * Uf
// (16 x 16, unit lower L) in shared memory (CTA 0 also stores it and the new diagonal block A[l0:l1, l0:l1] = Sinv),
// then F = Sinv * (W21 * A11)  (bw x l0), W21 (bw x l0, row-major, stride l0) gathered from the basis; also A21 = -F.
// One CTA per 16 columns j of A11: thread (ts, jj) accumulates rows t = ts (mod 16) of all 16 dot products of its
// column (A[t][j] is read once, coalesced over jj), the 16 slices are summed through shared memory.
__global__ void __launch_bounds__(256)
deim_inv_f_kernel(double* __restrict__ A, const double* __restrict__ W21, int k, int l0, int bw,
                  const double* __restrict__ bst, double* __restrict__ Sinv, double* __restrict__ F) {
    __shared__ double part[DEIM_BLOCK][DEIM_BLOCK][DEIM_BLOCK + 1];   // [slice][c][jj]
    __shared__ double e[DEIM_BLOCK][DEIM_BLOCK + 1], sinv[DEIM_BLOCK * DEIM_BLOCK];
    __shared__ double Ls[DEIM_BLOCK * DEIM_BLOCK], Zs[DEIM_BLOCK * DEIM_BLOCK], Li[DEIM_BLOCK * DEIM_BLOCK];
    constexpr int BB = DEIM_BLOCK * DEIM_BLOCK;
    const int tid = threadIdx.x, jj = tid % DEIM_BLOCK, ts = tid / DEIM_BLOCK, j = blockIdx.x * DEIM_BLOCK + jj;
    Ls[tid] = bst[BB + tid];
    Zs[tid] = bst[2 * BB + tid];
    Li[tid] = ts == jj ? 1.0 : 0.0;
    double acc[DEIM_BLOCK];
#pragma unroll
    for (int c = 0; c < DEIM_BLOCK; ++c) acc[c] = 0.0;
    if (j < l0)   // the long loads first: they overlap the small triangular work below
        for (int t = ts; t < l0; t += DEIM_BLOCK) {
            const double av = __ldcg(&A[(size_t)t * k + j]);
#pragma unroll
            for (int c = 0; c < DEIM_BLOCK; ++c) acc[c] = fma(c < bw ? __ldcg(&W21[(size_t)c * l0 + t]) : 0.0, av, acc[c]);
        }
#pragma unroll
    for (int c = 0; c < DEIM_BLOCK; ++c) part[ts][c][jj] = acc[c];
    __syncthreads();
    if (tid < bw) {   // column tid of inv(L) by forward substitution (L unit lower; strictly lower part stored)
        const int cc = tid;
        for (int r = cc + 1; r < bw; ++r) {
            double s = 0.0;
            for (int t = cc; t < r; ++t) s = fma(Ls[r * DEIM_BLOCK + t], Li[t * DEIM_BLOCK + cc], s);
            Li[r * DEIM_BLOCK + cc] = -s;
        }
    }
    {
        const int c = ts;   // thread (c, jj) sums the slices of E[c][j]
        double s = 0.0;
#pragma unroll
        for (int sl = 0; sl < DEIM_BLOCK; ++sl) s += part[sl][c][jj];
        e[c][jj] = s;
    }
    __syncthreads();
    {
        const int a = ts, b = jj;
        double s = 0.0;
        if (a < bw && b < bw)
            for (int t = (a > b ? a : b); t < bw; ++t) s = fma(Zs[a * DEIM_BLOCK + t], Li[t * DEIM_BLOCK + b], s);
        sinv[tid] = s;
        if (blockIdx.x == 0) {
            Sinv[tid] = s;
            if (a < bw && b < bw) A[(size_t)(l0 + a) * k + l0 + b] = s;
        }
    }
    __syncthreads();
    const int c = ts;
    if (j >= l0 || c >= bw) return;
    double s = 0.0;
#pragma unroll
    for (int c2 = 0; c2 < DEIM_BLOCK; ++c2) s = fma(sinv[c * DEIM_BLOCK + c2], e[c2][jj], s);
    F[(size_t)c * k + j] = s;
    A[(size_t)(l0 + c) * k + j] = -s;
}

// Panel end, step 3: A11 += C12 * F (rank-bw update), A12 = -C12 * Sinv.  One thread per entry of the l0 x (l0 + bw) block.
__global__ void __launch_bounds__(256)
deim_inv_update_kernel(double* __restrict__ A, int k, int l0, int bw, const double* __restrict__ C12,
                       const double* __restrict__ F, const double* __restrict__ Sinv) {
    const int s = blockIdx.x * 16 + (threadIdx.x & 15), t = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (t >= l0 || s >= l0 + bw) return;
    const double* c = C12 + t * DEIM_BLOCK;
    double acc = 0.0;
    if (s < l0) {
        for (int q = 0; q < bw; ++q) acc = fma(c[q], F[(size_t)q * k + s], acc);
        A[(size_t)t * k + s] += acc;
    } else {
        const int cc = s - l0;
        for (int q = 0; q < bw; ++q) acc = fma(c[q], Sinv[q * DEIM_BLOCK + cc], acc);
        A[(size_t)t * k + s] = -acc;
    }
}

// MOR_B200_DEIM_BLOCK=0 selects the streaming path (one pass over columns 1..l per greedy step, the
// reference's own access pattern); default is the blocked path whenever the basis layout allows it.
static bool deim_blocked_enabled() {
    const char* e = getenv("MOR_B200_DEIM_BLOCK");
    return !(e && e[0] == '0');
}
// Below this relative gap between the best and the second-best residual a greedy step of the BLOCKED path counts as
// a near-tie: its panel-wise arithmetic rounds differently from deim.jl:21-23, so the call is repeated with the
// streaming path (the reference's own formulation).  Default 32 k eps: what two summation orders of the k-term dot
// product of deim.jl:22 can differ by (measured: bench.py `deim.independent_check.residual_discrepancy`).
// MOR_B200_DEIM_TIE_FALLBACK=0 disables, MOR_B200_DEIM_TIE_TOL sets an absolute value.
static double deim_tie_tol(int64_t k) {
    const char* off = getenv("MOR_B200_DEIM_TIE_FALLBACK");
    if (off && off[0] == '0') return -1.0;
    const char* e = getenv("MOR_B200_DEIM_TIE_TOL");
    return e ? atof(e) : 32.0 * (double)k * 2.220446049250313e-16;
}

// Launch a step kernel; `pdl`: the previous operation on the stream was a step kernel too, so this grid may be
// scheduled as soon as that one's CTAs have exited or reached their tail (programmatic stream serialization).
template <class... KArgs, class... Args>
static cudaError_t launch_step(void (*kern)(KArgs...), int grid, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(DEIM_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static int32_t deim_run(const double* B, int64_t m, int64_t k, int64_t ld, int64_t* idx_host, const cudaEvent_t* col_ready,
                        int group_cols, bool allow_blocked) {
    Ctx& cx = ctx();
    for (double& t : cx.timings) t = 0.0;

    std::vector<int64_t> all_m;
    MOR_TRY(comm_allgather_i64(m, all_m));
    int64_t row_offset = 0, m_global = 0;
    for (int r = 0; r < cx.nranks; ++r) {
        if (r < cx.rank) row_offset += all_m[r];
        m_global += all_m[r];
    }
    MOR_ARG(m_global >= 1, "deim_interpolation_indices: empty basis");
    MOR_ARG(cx.nranks <= DEIM_MAXRANKS, "deim_interpolation_indices: more than 8 ranks");

    constexpr int R = 4;
    const int64_t CH = 2 * R * DEIM_THREADS;
    int64_t nchunks = (m + CH - 1) / CH;
    int grid = (int)(nchunks < (int64_t)cx.num_sms * 8 ? nchunks : (int64_t)cx.num_sms * 8);
    if (grid < 1) grid = 1;
    const bool vec = ((uintptr_t)B % 16 == 0) && (ld % 2 == 0);

    // Blocked path (k > 16): the residuals of 16 columns at a time with respect to all earlier points come
    // from ONE pass over the earlier columns (deim_panel_kernel); the greedy steps inside a block run the
    // same streaming kernel on that 16-column residual panel -- about 8 m k (k / 32 + 18) bytes of traffic
    // instead of 4 m k (k + 1).  It needs a 16-byte aligned basis with an even leading dimension (bulk
    // copies) and 160 m bytes of workspace; every rank must take the same path (the gathered records differ).
    const int64_t ldp = (m + 1) & ~int64_t(1);
    bool blocked = allow_blocked && deim_blocked_enabled() && k > DEIM_BLOCK && (m == 0 || (vec && ld >= ldp));
    DevBuf panel;
    if (blocked && m > 0 && panel.alloc(sizeof(double) * (size_t)ldp * (DEIM_BLOCK + DEIM_SUB)) != MOR_OK) {
        cudaGetLastError();   // no room for the panel: the streaming path needs no workspace
        blocked = false;
    }
    if (cx.nranks > 1) {   // all ranks or none
        std::vector<int64_t> votes;
        MOR_TRY(comm_allgather_i64(blocked ? 1 : 0, votes));
        for (int64_t v : votes) blocked = blocked && v != 0;
        if (!blocked) panel.release();
    }
    MOR_TRY(deim_mailbox());            // own mailbox (+ peers' over NVLink when every rank could map them)
    const bool fused = cx.nranks == 1 || cx.peers_ok;

    const size_t kk = (size_t)k * k;
    // exchanged record: value, index, then -- streaming path: the winner's row of the basis (k); blocked path: its rows
    // of the panel (16) and of the sub-panel (4) -- then the runner-up value
    const int krec = blocked ? 0 : (int)k;
    const int rec = krec + 3 + (blocked ? DEIM_BLOCK + DEIM_SUB : 0);
    DevBuf state, cands;
    // state: blocked: A (k*k), F (16 k), C12 (16 k), gathered W21 | W12 (32 k), packed Cp, Sinv, bst, sst, Dsub;
    //        streaming: W, Uf, L, Z, Zt (k*k each), c (k); both: margins, values (k), idx (k int64), status
    const size_t cp_doubles = (size_t)((k + DEIM_BLOCK - 1) / DEIM_BLOCK) * DEIM_BLOCK * 20;
    const size_t small_doubles = DEIM_BLOCK * DEIM_BLOCK + BST_DOUBLES + SST_DOUBLES + DEIM_BLOCK * DEIM_SUB;
    // every piece starts on a 128-byte boundary (the panel kernel fetches the packed coefficients with bulk copies)
    auto r16 = [](size_t nd) { return (nd + 15) & ~size_t(15); };
    const size_t kb16 = r16((size_t)k * DEIM_BLOCK);
    const size_t ndoubles = (blocked ? r16(kk) + 4 * kb16 + r16(cp_doubles) + r16(small_doubles) + 64 : 5 * r16(kk) + r16((size_t)k)) + 2 * r16((size_t)k);
    MOR_TRY(state.alloc(sizeof(double) * ndoubles + sizeof(long long) * (size_t)k + 64));
    MOR_TRY(cands.alloc(sizeof(Cand) * (size_t)grid));
    MOR_CUDA(cudaMemsetAsync(state.p, 0, state.bytes, cx.stream));
    double* p = state.as<double>();
    auto take = [&](size_t nd) {
        double* q = p;
        p += r16(nd);
        return q;
    };
    double *W = nullptr, *A = nullptr, *Fm = nullptr, *C12 = nullptr, *W21 = nullptr, *W12 = nullptr, *Cp = nullptr, *Sinv = nullptr,
           *bst = nullptr, *sst = nullptr, *Dsub = nullptr, *Uf = nullptr, *Lm = nullptr, *Z = nullptr, *Zt = nullptr, *c = nullptr;
    if (blocked) {
        A = take(kk);
        Fm = take((size_t)k * DEIM_BLOCK);
        C12 = take((size_t)k * DEIM_BLOCK);
        W21 = take((size_t)k * DEIM_BLOCK);      // W21 and W12 are contiguous (kb16 doubles each): one all-reduce
        W12 = take((size_t)k * DEIM_BLOCK);
        Cp = take(cp_doubles);
        Sinv = take(DEIM_BLOCK * DEIM_BLOCK);
        bst = take(BST_DOUBLES);
        sst = take(SST_DOUBLES);
        Dsub = take(DEIM_BLOCK * DEIM_SUB);
    } else {
        W = take(kk);
        Uf = take(kk);
        Lm = take(kk);
        Z = take(kk);
        Zt = take(kk);
        c = take((size_t)k);
    }
    double* margins_d = take((size_t)k);
    double* values_d = take((size_t)k);
    long long* idx_d = reinterpret_cast<long long*>(p);
    int* status_d = reinterpret_cast<int*>(idx_d + k);
    unsigned long long* tail_ns_d = reinterpret_cast<unsigned long long*>(status_d + 2);
    double* R2 = blocked && m > 0 ? panel.as<double>() + (size_t)ldp * DEIM_BLOCK : nullptr;
    // pipelined ingest: later columns of the basis are not there yet, so a record carries the winner's row only up
    // to the current panel and the rest is gathered at each panel boundary
    const bool defer = blocked && col_ready != nullptr;
    if (!blocked && col_ready != nullptr)
        for (int g = 0; g * group_cols < (int)k; ++g) MOR_CUDA(cudaStreamWaitEvent(cx.stream, col_ready[g], 0));

    TailArgs ta{};
    ta.ticket = cx.box_ticket;
    ta.cands = cands.as<Cand>();
    ta.U = B;
    ta.ldU = ld;
    ta.k = (int)k;
    ta.krec = krec;
    ta.row_offset = row_offset;
    ta.ll = cx.box_ll;
    for (int r = 0; r < cx.nranks; ++r) ta.peer_ll[r] = fused ? cx.peer_ll[r] : cx.box_ll;
    ta.send = cx.box + (size_t)DEIM_SLOTS * DEIM_MAXRANKS * DEIM_REC_MAX;
    ta.rank = cx.rank;
    ta.nranks = cx.nranks;
    ta.rec = rec;
    ta.mode = fused ? 0 : 1;
    ta.W = W;
    ta.idx_out = idx_d;
    ta.margins = margins_d;
    ta.values = values_d;
    ta.blocked = blocked ? 1 : 0;
    ta.bst = bst;
    ta.sst = sst;
    ta.Dsub = Dsub;
    ta.status = status_d;
    ta.tail_ns = tail_ns_d;

    static const bool use_pdl = !(getenv("MOR_B200_DEIM_PDL") && getenv("MOR_B200_DEIM_PDL")[0] == '0');
    Timer t_total(MOR_T_TOTAL), t_panel(MOR_T_DEIM_PANEL), t_upd(MOR_T_DEIM_UPDATE);
    t_total.start();
    for (int i = 0; i < (int)k; ++i) {
        const int l0 = blocked ? (i / DEIM_BLOCK) * DEIM_BLOCK : 0;   // first column of the current block
        const int j = i - l0;                                          // step inside the block
        const int bw = blocked ? (int)(k - l0 < DEIM_BLOCK ? k - l0 : DEIM_BLOCK) : (int)k;
        // the matrix the streaming kernel works on: the basis itself, or the residual panel of this block
        const double* Pn = (blocked && l0 > 0) ? panel.as<double>() : B + (size_t)l0 * ld;
        const int64_t ldn = (blocked && l0 > 0) ? ldp : ld;
        // second level: sub-panels of 4 columns inside the panel
        const int s0 = blocked ? (j / DEIM_SUB) * DEIM_SUB : 0;       // first panel column of the current sub-panel
        const int jj = j - s0;                                        // step inside the sub-panel
        const int sw = blocked ? (bw - s0 < DEIM_SUB ? bw - s0 : DEIM_SUB) : 0;
        const double* P2 = (blocked && s0 > 0) ? R2 : Pn + (size_t)s0 * ldn;
        const int64_t ld2 = (blocked && s0 > 0) ? ldp : ldn;
        // pipelined host ingest: columns [i, i + group_cols) are first touched by this step (the panel of a
        // block reads columns < l0 + 16 and DEIM_BLOCK == group_cols; the streaming step i reads columns <= i)
        if (defer && i == 0) MOR_CUDA(cudaStreamWaitEvent(cx.stream, col_ready[0], 0));
        if (blocked && j == 0 && l0 > 0) {
            t_upd.start();
            deim_coef_kernel<<<(l0 + 7) / 8, 256, 0, cx.stream>>>(A, (int)k, l0, bw, W12, DEIM_BLOCK, Cp, C12);
            count_launch();
            t_upd.stop();
            t_panel.start();
            if (m > 0) MOR_TRY(deim_panel(B, ld, l0, bw, Cp, m, panel.as<double>(), ldp));
            t_panel.stop();
        }
        ta.i = i;
        ta.Pn = blocked && m > 0 ? Pn : nullptr;
        ta.ldn = ldn;
        ta.bw = bw;
        ta.P2 = blocked && m > 0 ? P2 : nullptr;
        ta.ld2 = ld2;
        ta.sw = sw;
        ta.j = j;
        ta.jj = jj;
        ta.s0 = s0;
        ta.last = i == (int)k - 1 ? 1 : 0;
        const unsigned long long seq = cx.deim_seq + (unsigned long long)i;
        ta.slot = (int)(seq % DEIM_SLOTS);
        ta.tag = (unsigned)(seq % 0xfffffff0ull) + 1u;   // never 0 (a zeroed mailbox matches no step)
        const bool pdl = use_pdl && blocked && fused && j > 0;   // the stream's previous operation was a step kernel
        if (blocked && s0 > 0 && jj == 0) {
            // residuals of the next 4 panel columns in one pass over the s0 earlier ones, fused with this step's argmax
            MOR_CUDA(launch_step(deim_step4_kernel<R>, grid, cx.stream, pdl, Pn, m, ldn, s0, sw, (const double*)Dsub, R2, ldp,
                                 row_offset, ta));
        } else if (blocked) {
            MOR_CUDA(launch_step(deim_step_kernel<R, true>, grid, cx.stream, pdl, P2, m, ld2, jj,
                                 (const double*)(sst + 3 * DEIM_SUB * DEIM_SUB), row_offset, ta));
        } else if (vec) {
            deim_step_kernel<R, true><<<grid, DEIM_THREADS, 0, cx.stream>>>(B, m, ld, i, c, row_offset, ta);
        } else {
            deim_step_kernel<R, false><<<grid, DEIM_THREADS, 0, cx.stream>>>(B, m, ld, i, c, row_offset, ta);
        }
        count_launch();
        if (!fused) {   // NCCL fallback: all-gather the records, then the global part of the tail as its own launch
            double* gathered = cx.box + (size_t)ta.slot * cx.nranks * rec;
            MOR_TRY(comm_allgather(ta.send, gathered, sizeof(double) * (size_t)rec));
            deim_tail_global_kernel<<<1, DEIM_THREADS, 0, cx.stream>>>(ta, gathered);
            count_launch();
        }
        if (!blocked) {
            deim_update_kernel<<<1, 1024, 0, cx.stream>>>(W, Uf, Lm, Z, Zt, c, i, (int)k, ta.last, status_d);
            count_launch();
        } else if (j == bw - 1 && l0 + bw < (int)k) {
            // The panel is complete.  Fetch the two pieces of P'U that are needed now (the rows of the new points in
            // the earlier columns, the rows of all points in the NEXT panel's columns), then extend A = inv(P'U) by
            // the panel's bw rows / columns (block-inverse formula with the panel's own LU factors).
            t_upd.start();
            const int l1 = l0 + bw, bwn = (int)(k - l1 < DEIM_BLOCK ? k - l1 : DEIM_BLOCK);
            // pipelined host ingest: the next panel's columns must have landed before their rows are gathered
            if (defer) MOR_CUDA(cudaStreamWaitEvent(cx.stream, col_ready[l1 / group_cols], 0));
            // W21 (bw x l0) and W12 (l1 x 16) sit back to back in one buffer: one launch, one all-reduce
            W12 = W21 + r16((size_t)bw * l0);
            deim_gather_kernel<<<(bw * l0 + l1 * DEIM_BLOCK + 255) / 256, 256, 0, cx.stream>>>(B, m, ld, row_offset, idx_d, l0, bw, bwn,
                                                                                           W21, W12);
            if (cx.nranks > 1) MOR_TRY(comm_allreduce_sum(W21, r16((size_t)bw * l0) + (size_t)l1 * DEIM_BLOCK));
            deim_inv_f_kernel<<<l0 > 0 ? (l0 + DEIM_BLOCK - 1) / DEIM_BLOCK : 1, 256, 0, cx.stream>>>(A, W21, (int)k, l0, bw, bst, Sinv, Fm);
            if (l0 > 0) {
                deim_inv_update_kernel<<<dim3((l0 + bw + 15) / 16, (l0 + 15) / 16), 256, 0, cx.stream>>>(A, (int)k, l0, bw, C12,
                                                                                                   Fm, Sinv);
                count_launch();
            }
            count_launch();
            count_launch();
            t_upd.stop();
        }
    }
    t_total.stop();
    if (cx.nranks > 1) cx.deim_seq += (unsigned long long)k;   // a one-rank (solo) call must not desynchronise the ranks' step counters
    MOR_CUDA(cudaGetLastError());
    std::vector<long long> idx_h((size_t)k);
    cx.deim_margins.assign((size_t)k, 0.0);
    cx.deim_values.assign((size_t)k, 0.0);
    int status_h = 0;
    unsigned long long tail_ns = 0;
    MOR_CUDA(cudaMemcpyAsync(idx_h.data(), idx_d, sizeof(long long) * (size_t)k, cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaMemcpyAsync(cx.deim_margins.data(), margins_d, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaMemcpyAsync(cx.deim_values.data(), values_d, sizeof(double) * (size_t)k, cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaMemcpyAsync(&status_h, status_d, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaMemcpyAsync(&tail_ns, tail_ns_d, sizeof(tail_ns), cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaStreamSynchronize(cx.stream));
    t_total.collect();
    t_panel.collect();
    t_upd.collect();
    cx.timings[MOR_T_DEIM_TAIL] = (double)tail_ns * 1e-6;
    cx.timings[MOR_T_DEIM_STEP] = cx.timings[MOR_T_TOTAL] - cx.timings[MOR_T_DEIM_PANEL] - cx.timings[MOR_T_DEIM_UPDATE];
    cx.timings[MOR_T_DEIM_BLOCKED] = blocked ? 1.0 : 0.0;
    cx.timings[MOR_T_DEIM_EXCHANGE] = cx.nranks == 1 ? 0.0 : (fused ? 1.0 : 2.0);
    for (int64_t i = 0; i < k; ++i) idx_host[i] = (int64_t)idx_h[(size_t)i];
    if (status_h < 0) {
        set_error("deim_interpolation_indices: the peer exchange timed out (a rank of the job did not reach the same greedy step)");
        return MOR_ERR_CUDA;
    }
    if (status_h != 0) {
        set_error("deim_interpolation_indices: P'U is exactly singular at step " + std::to_string(status_h + 1) +
                  " (SingularException in the reference)");
        return MOR_ERR_SINGULAR;
    }
    return MOR_OK;
}

int32_t deim_indices_device(const double* B, int64_t m, int64_t k, int64_t ld, int64_t* idx_host,
                            const cudaEvent_t* col_ready, int group_cols) {
    Ctx& cx = ctx();
    MOR_ARG(B != nullptr && idx_host != nullptr, "deim_interpolation_indices: null pointer");
    MOR_ARG(k >= 1 && k <= DEIM_MAXK, "deim_interpolation_indices: number of basis columns must be in 1..1024");
    MOR_ARG(m >= 0 && ld >= (m > 0 ? m : 1), "deim_interpolation_indices: bad m / ld");
    MOR_TRY(deim_run(B, m, k, ld, idx_host, col_ready, group_cols, true));
    cx.timings[MOR_T_DEIM_FALLBACK] = 0.0;
    // Near-tie guard: step 1 is the argmax of |u_1| itself (no arithmetic, the same in every formulation); from
    // step 2 on, a decision of the blocked path closer than the tolerance is re-made in the reference's formulation.
    const double tol = deim_tie_tol(k);
    if (cx.timings[MOR_T_DEIM_BLOCKED] != 0.0 && tol >= 0.0) {
        double mn = 2.0;
        for (size_t i = 1; i < cx.deim_margins.size(); ++i) mn = cx.deim_margins[i] < mn ? cx.deim_margins[i] : mn;
        if (mn < tol) {
            const double first_ms = cx.timings[MOR_T_TOTAL];
            if (col_ready != nullptr)   // the whole basis must have landed before the streaming path starts
                for (int g = 0; g * group_cols < (int)k; ++g) MOR_CUDA(cudaStreamWaitEvent(cx.stream, col_ready[g], 0));
            MOR_TRY(deim_run(B, m, k, ld, idx_host, nullptr, group_cols, false));
            cx.timings[MOR_T_TOTAL] += first_ms;
            cx.timings[MOR_T_DEIM_FALLBACK] = 1.0;
        }
    }
    return MOR_OK;
}

// DEIM projector of the reduced nonlinear term, reference src/deim.jl:150:
//     projector = ((@view U[indices, :])' \ (U' * V))'        (pod_dim x deim_dim)
// U'V is a tall-skinny contraction over the row shards (gemm_tn + all-reduce), U[indices, :] is
// gathered from whichever rank owns each selected row, the kd x kd solve is LU with partial
// pivoting (Julia's `\` for a general square matrix), replicated on every rank.
int32_t deim_projector_device(const double* U, int64_t m, int kd, int64_t ldu, const double* V, int kp, int64_t ldv,
                              const int64_t* idx_host, double* P_dev, int64_t ldp) {
    Ctx& cx = ctx();
    MOR_ARG(U && V && idx_host && P_dev, "deim_projector: null pointer");
    MOR_ARG(kd >= 1 && kd <= 1024 && kp >= 1 && kp <= 4096 && ldp >= kp, "deim_projector: bad dimensions");
    std::vector<int64_t> all_m;
    MOR_TRY(comm_allgather_i64(m, all_m));
    int64_t row_offset = 0, m_global = 0;
    for (int r = 0; r < cx.nranks; ++r) {
        if (r < cx.rank) row_offset += all_m[r];
        m_global += all_m[r];
    }
    for (int i = 0; i < kd; ++i)
        MOR_ARG(idx_host[i] >= 1 && idx_host[i] <= m_global, "deim_projector: interpolation index out of range");
    DevBuf UtV, PtU, d_idx, d_fail;
    MOR_TRY(UtV.alloc(sizeof(double) * (size_t)kd * kp));
    MOR_TRY(PtU.alloc(sizeof(double) * (size_t)kd * kd));
    MOR_TRY(d_idx.alloc(sizeof(long long) * (size_t)kd));
    MOR_TRY(d_fail.alloc(sizeof(int)));
    std::vector<long long> idx_ll(idx_host, idx_host + kd);
    MOR_CUDA(cudaMemcpyAsync(d_idx.p, idx_ll.data(), sizeof(long long) * (size_t)kd, cudaMemcpyHostToDevice, cx.stream));
    for (double& t : cx.timings) t = 0.0;
    Timer t_proj(MOR_T_PROJ);
    t_proj.start();
    MOR_TRY(gemm_tn(U, ldu, kd, V, ldv, kp, m, UtV.as<double>(), kd));
    t_proj.stop();
    MOR_TRY(gather_rows(U, m, ldu, row_offset, d_idx.as<long long>(), kd, PtU.as<double>(), kd));
    if (cx.nranks > 1) {
        MOR_TRY(comm_allreduce_sum(UtV.as<double>(), (size_t)kd * kp));
        MOR_TRY(comm_allreduce_sum(PtU.as<double>(), (size_t)kd * kd));
    }
    MOR_TRY(small_lu_solve(PtU.as<double>(), kd, kd, /*transpose=*/true, UtV.as<double>(), kp, kd, d_fail.as<int>()));
    MOR_TRY(transpose(UtV.as<double>(), kd, kp, kd, P_dev, ldp));
    int fail = 0;
    MOR_CUDA(cudaMemcpyAsync(&fail, d_fail.p, sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    MOR_CUDA(cudaStreamSynchronize(cx.stream));
    t_proj.collect();
    if (fail != 0) {
        set_error("deim_projector: U[indices, :] is exactly singular (SingularException in the reference)");
        return MOR_ERR_SINGULAR;
    }
    return MOR_OK;
}

}  // namespace mor

using namespace mor;

extern "C" {

int32_t mor_b200_deim_projector_dev(mor_mat_t U, mor_mat_t V, const int64_t* idx, mor_mat_t P_out) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_deim_projector_dev(U, V, idx, P_out));
    MOR_ARG(U && V && idx && P_out, "mor_b200_deim_projector_dev: null argument");
    MOR_ARG(U->m == V->m && P_out->m == V->n && P_out->n == U->n, "mor_b200_deim_projector_dev: P_out must be pod_dim x deim_dim");
    MOR_TRY(require_ctx());
    return deim_projector_device(U->p, U->m, (int)U->n, U->ld, V->p, (int)V->n, V->ld, idx, P_out->p, P_out->ld);
}

int32_t mor_b200_deim_projector(const double* U, int64_t m, int64_t kd, int64_t ldu, const double* V, int64_t kp,
                                int64_t ldv, const int64_t* idx, double* P_out, int64_t ldp) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_deim_projector(U, m, kd, ldu, V, kp, ldv, idx, P_out, ldp);
    MOR_ARG(U && V && idx && P_out, "mor_b200_deim_projector: null pointer");
    MOR_ARG(m >= 0 && kd >= 1 && kp >= 1 && ldu >= (m > 0 ? m : 1) && ldv >= (m > 0 ? m : 1) && ldp >= kp,
            "mor_b200_deim_projector: bad dimensions");
    MOR_TRY(require_ctx());
    mor_mat_t dU = nullptr, dV = nullptr, dP = nullptr;
    int32_t st = mor_b200_mat_alloc(m, kd, &dU);
    if (st == MOR_OK) st = mor_b200_mat_alloc(m, kp, &dV);
    if (st == MOR_OK) st = mor_b200_mat_alloc(kp, kd, &dP);
    if (st == MOR_OK) st = mor_b200_mat_upload(dU, U, ldu);
    if (st == MOR_OK) st = mor_b200_mat_upload(dV, V, ldv);
    if (st == MOR_OK) st = mor_b200_deim_projector_dev(dU, dV, idx, dP);
    if (st == MOR_OK) st = mor_b200_mat_download(dP, P_out, ldp);
    mor_b200_mat_free(dU);
    mor_b200_mat_free(dV);
    mor_b200_mat_free(dP);
    return st;
}

int32_t mor_b200_deim_last_margins(double* margins, double* values, int64_t cap, int64_t* count, double* min_margin,
                                   int64_t* min_step) {
    MOR_API_LOCK;
    const std::vector<double>& mg = ctx().deim_margins;
    const std::vector<double>& vl = ctx().deim_values;
    if (count) *count = (int64_t)mg.size();
    for (int64_t i = 0; i < cap && i < (int64_t)mg.size(); ++i) {
        if (margins) margins[i] = mg[(size_t)i];
        if (values) values[i] = vl[(size_t)i];
    }
    double mn = 1.0;
    int64_t at = 0;
    for (size_t i = 1; i < mg.size(); ++i)
        if (at == 0 || mg[i] < mn) {
            mn = mg[i];
            at = (int64_t)i + 1;
        }
    if (min_margin) *min_margin = mn;
    if (min_step) *min_step = at;
    return MOR_OK;
}

int32_t mor_b200_deim_indices_dev(mor_mat_t B, int64_t* idx_out) {
    MOR_API_LOCK;
    MOR_ROUTE_SOLO(mor_b200_deim_indices_dev(B, idx_out));
    MOR_ARG(B != nullptr, "mor_b200_deim_indices_dev: null matrix");
    MOR_TRY(require_ctx());
    return deim_indices_device(B->p, B->m, B->n, B->ld, idx_out);
}

int32_t mor_b200_deim_indices(const double* B, int64_t m, int64_t k, int64_t ldb, int64_t* idx_out) {
    MOR_API_LOCK;
    if (multi_on() && !is_worker()) return multi_deim_indices(B, m, k, ldb, idx_out);
    MOR_ARG(B != nullptr && idx_out != nullptr, "mor_b200_deim_indices: null pointer");
    MOR_ARG(m >= 0 && k >= 1 && ldb >= (m > 0 ? m : 1), "mor_b200_deim_indices: bad dimensions");
    MOR_TRY(require_ctx());
    Ctx& cx = ctx();
    // device copy of the basis: large ones reuse the parked ingest buffer (a cudaMalloc / cudaFree pair of tens of GB
    // costs ~0.1 s and serialises the GPUs of a process)
    struct Dev { double* p; int64_t m, n, ld; } dv{nullptr, m, k, std::max<int64_t>((m + 1) & ~int64_t(1), 2)};
    DevBuf dbuf;
    MOR_TRY(ingest_alloc(dbuf, sizeof(double) * (size_t)dv.ld * (size_t)k));
    dv.p = dbuf.as<double>();
    if (dv.ld != m) MOR_CUDA(cudaMemset2DAsync(dv.p + m, (size_t)dv.ld * 8, 0, 8, (size_t)k, cx.stream));   // keep the pad row finite
    Dev* d = &dv;
    // Pipelined ingest: the basis travels in groups of 16 columns on an upload stream; the greedy loop touches
    // columns left to right (a panel needs the columns up to its own), so it starts on group 0 while the rest is
    // still on the PCIe link -- the 34 GB of a 16M x 256 basis hide all but the last panel of the computation.
    constexpr int GROUP = DEIM_BLOCK;
    const int ngroups = (int)((k + GROUP - 1) / GROUP);
    std::vector<cudaEvent_t> ev((size_t)ngroups + 2, nullptr);   // [ngroups] = start of the upload, [ngroups+1] = pad row zeroed
    int32_t st = MOR_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (st == MOR_OK && e != cudaSuccess) st = cuda_fail(e, what, __FILE__, __LINE__);
    };
    if (!cx.aux_stream) cu(cudaStreamCreateWithFlags(&cx.aux_stream, cudaStreamNonBlocking), "cudaStreamCreate(aux)");
    for (auto& e : ev) cu(cudaEventCreate(&e), "cudaEventCreate");
    if (st == MOR_OK) {
        cu(cudaEventRecord(ev[(size_t)ngroups + 1], cx.stream), "cudaEventRecord");   // the pad row is zeroed on cx.stream
        cu(cudaStreamWaitEvent(cx.aux_stream, ev[(size_t)ngroups + 1], 0), "cudaStreamWaitEvent");
        cu(cudaEventRecord(ev[(size_t)ngroups], cx.aux_stream), "cudaEventRecord");
        for (int g = 0; g < ngroups && st == MOR_OK; ++g) {
            const int64_t c0 = (int64_t)g * GROUP, nc = k - c0 < GROUP ? k - c0 : GROUP;
            if (m > 0 && st == MOR_OK) st = h2d_2d(d->p + (size_t)c0 * d->ld, d->ld, B + (size_t)c0 * ldb, ldb, m, nc, cx.aux_stream);
            cu(cudaEventRecord(ev[(size_t)g], cx.aux_stream), "cudaEventRecord");
        }
    }
    if (st == MOR_OK) st = deim_indices_device(d->p, d->m, d->n, d->ld, idx_out, ev.data(), GROUP);
    if (cx.aux_stream) cudaStreamSynchronize(cx.aux_stream);     // nothing may still read the caller's buffer
    if (st == MOR_OK) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[(size_t)ngroups], ev[(size_t)ngroups - 1]) == cudaSuccess) cx.timings[MOR_T_H2D] = ms;
    }
    for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    cudaStreamSynchronize(cx.stream);
    return st;
}

}  // extern "C"
