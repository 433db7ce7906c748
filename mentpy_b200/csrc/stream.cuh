// Streaming regime: ONE large window (2^w amplitudes resident in HBM, w up to ~33 per GPU), one
// angle set; every pass streams the state once, in place, and applies up to 5 consecutive
// measurements to it (np_simulator_sv.py:164-225 per measurement).
//
// Per pass with K fused measurements on physical slots s_0..s_{K-1}: a thread owns the 2^K
// amplitudes that differ only in those slots (its "group"), loads them with 128-bit loads,
// applies the K pair-reductions + CZ signs in registers and stores the survivors -- one read and
// one write of the state per K measurements (HBM-bound: 32 B per live amplitude per pass).
// Neighbouring threads own neighbouring groups, so every load/store instruction of a warp covers
// consecutive 16-byte elements whenever the fused slots are above bit 4.
//
// Dead slots (tail of the pattern: measured, nothing appended) are squeezed out of the thread
// index with zero-field inserts, so the tail passes touch only the live 2^n amplitudes.
// Normalisation is deferred: passes apply an exact power-of-two rescale; the true norm is taken
// once at gather time.  (1 + e^{i theta}) phase factors are accumulated on the host.
#pragma once
#include "common.cuh"

namespace mbqc {

__device__ __forceinline__ uint64_t insert_zero_field(uint64_t g, uint32_t pos, uint32_t width) {
    const uint64_t lo = g & ((1ull << pos) - 1ull);
    return ((g >> pos) << (pos + width)) | lo;
}

// window seed in kernel-friendly form: input gather + initial CZ signs grouped by slot distance d,
// sign = parity( XOR_d  g & (g >> d) & pair_mask[d] )  (1 distance for a linear cluster, {1, rows}
// for a grid) instead of a loop over all window bits per amplitude
struct SeedDev {
    const double2* input;  // null: |+> inputs
    int32_t n_in;
    int32_t n_dist;
    double scale;
    int32_t in_slot[kMaxIO];
    int32_t dist[MBQC_MAX_WINDOW];
    uint64_t pair_mask[MBQC_MAX_WINDOW];
    // seeded passes (|+> inputs only): sign of group element l relative to the group's base index
    //   sign(l) = sign(base) ^ XOR_{j in l} parity(base & nbr[j]) ^ bit l of pair_parity
    uint64_t nbr[MBQC_STREAM_MAX_FUSE];  // initial-CZ neighbours of fused slot j (both directions)
    uint32_t pair_parity;                // bit l: parity of the initial CZ edges inside subset l
    uint32_t in_bit[MBQC_STREAM_MAX_FUSE];  // input-index bit fed by fused slot j (0: not an input)
};

__device__ __forceinline__ double2 seed_amplitude(const SeedDev& p, uint64_t g) {
    double2 v = make_double2(p.scale, 0.0);
    if (p.input) {
        uint32_t src = 0;
        for (int q = 0; q < p.n_in; ++q) src |= (uint32_t)((g >> p.in_slot[q]) & 1ull) << (p.n_in - 1 - q);
        v = __ldg(p.input + src);
        v.x *= p.scale;
        v.y *= p.scale;
    }
    uint64_t acc = 0;
    for (int q = 0; q < p.n_dist; ++q) acc ^= g & (g >> p.dist[q]) & p.pair_mask[q];
    if (__popcll(acc) & 1) {
        v.x = -v.x;
        v.y = -v.y;
    }
    return v;
}

template <int K, bool SEED>
__global__ void __launch_bounds__(256) stream_steps_kernel(double2* __restrict__ state,
                                                           const __grid_constant__ mbqc_stream_desc d,
                                                           const __grid_constant__ SeedDev seed) {
    constexpr int N = 1 << K;
    uint64_t ofs[N];
#pragma unroll
    for (int l = 0; l < N; ++l) {
        uint64_t o = 0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (l & (1 << j)) o += d.elem_offset[j];
        ofs[l] = o;
    }
    uint32_t dead_local = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
        if (!((d.append_mask >> j) & 1u)) dead_local |= 1u << j;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < d.n_groups; t += stride) {
        uint64_t g = t;
        for (int r = 0; r < d.n_ranges; ++r) g = insert_zero_field(g, d.range_pos[r], d.range_width[r]);
        double2* base = state + g;
        double re[N], im[N];
        if constexpr (!SEED) {
#pragma unroll
            for (int l = 0; l < N; ++l) {
                const double2 v = base[ofs[l]];
                re[l] = v.x * d.scale;
                im[l] = v.y * d.scale;
            }
        }
        const uint64_t gfull = g | d.index_or;
        if constexpr (SEED) {  // first pass of a pattern: generate the |+>^w seed instead of reading it
            uint64_t acc = 0;
            for (int q = 0; q < seed.n_dist; ++q) acc ^= gfull & (gfull >> seed.dist[q]) & seed.pair_mask[q];
            uint32_t pj = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) pj |= ((uint32_t)__popcll(gfull & seed.nbr[j]) & 1u) << j;
            const uint32_t par0 = (uint32_t)__popcll(acc) & 1u;
            const double a = seed.scale * d.scale;
            uint32_t src0 = 0;
            if (seed.input)
                for (int q = 0; q < seed.n_in; ++q) src0 |= (uint32_t)((gfull >> seed.in_slot[q]) & 1ull) << (seed.n_in - 1 - q);
#pragma unroll
            for (int l = 0; l < N; ++l) {
                const uint32_t sg = (par0 ^ ((uint32_t)__popc((uint32_t)l & pj) & 1u) ^ ((seed.pair_parity >> l) & 1u)) << 31;
                if (seed.input) {
                    uint32_t src = src0;
#pragma unroll
                    for (int j = 0; j < K; ++j)
                        if (l & (1 << j)) src |= seed.in_bit[j];
                    const double2 v = __ldg(seed.input + src);
                    re[l] = flip_sign(v.x * a, sg);
                    im[l] = flip_sign(v.y * a, sg);
                } else {
                    re[l] = flip_sign(a, sg);
                    im[l] = 0.0;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double c = d.cos_t[j], s = d.sin_t[j];
            const uint32_t pg = (uint32_t)__popcll(gfull & d.nbr_mask[j]) & 1u;
            const uint32_t ml = d.local_mask[j];
#pragma unroll
            for (int l = 0; l < N; ++l) {
                if (l & (1 << j)) continue;
                const int lj = l | (1 << j);
                const double tr = fma(c, re[lj], fma(s, im[lj], re[l]));
                const double ti = fma(c, im[lj], fma(-s, re[lj], im[l]));
                re[l] = tr;
                im[l] = ti;
                const uint32_t sg = (pg ^ ((uint32_t)__popc((uint32_t)l & ml) & 1u)) << 31;
                re[lj] = flip_sign(tr, sg);
                im[lj] = flip_sign(ti, sg);
            }
        }
#pragma unroll
        for (int l = 0; l < N; ++l)
            if (((uint32_t)l & dead_local) == 0) base[ofs[l]] = make_double2(re[l], im[l]);
    }
}

// Low-slot variant: every fused slot is one of the 5 lane bits of the index (slot < 5).  The
// register-group kernel above would make each thread walk 2^K neighbouring elements (lanes 2^K * 16
// bytes apart: uncoalesced); here a thread owns ONE element, a warp one contiguous 512-byte span,
// and the pair partner is fetched with __shfl_xor -- the in-warp form of the "high strides through
// shared memory" staging.  Descriptor convention: ranges squeeze out dead slots only (all >= 5),
// n_groups = number of live local elements (a multiple of 32).
__global__ void __launch_bounds__(256) stream_lane_kernel(double2* __restrict__ state,
                                                          const __grid_constant__ mbqc_stream_desc d) {
    const int K = d.n_fused;
    uint32_t dead_mask = 0;
    for (int j = 0; j < K; ++j)
        if (!((d.append_mask >> j) & 1u)) dead_mask |= (uint32_t)d.elem_bit[j];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < d.n_groups; t += stride) {
        uint64_t i = t;
        for (int r = 0; r < d.n_ranges; ++r) i = insert_zero_field(i, d.range_pos[r], d.range_width[r]);
        double2 v = state[i];
        v.x *= d.scale;
        v.y *= d.scale;
        const uint64_t gfull = i | d.index_or;
#pragma unroll
        for (int j = 0; j < MBQC_STREAM_MAX_FUSE; ++j) {
            if (j < K) {
                const uint32_t bit = (uint32_t)d.elem_bit[j];
                const double px = __shfl_xor_sync(0xffffffffu, v.x, (int)bit);
                const double py = __shfl_xor_sync(0xffffffffu, v.y, (int)bit);
                const bool hi = ((uint32_t)i & bit) != 0;
                const double a0x = hi ? px : v.x, a0y = hi ? py : v.y;  // bit-0 partner
                const double a1x = hi ? v.x : px, a1y = hi ? v.y : py;  // bit-1 partner
                const double c = d.cos_t[j], s = d.sin_t[j];
                double tr = fma(c, a1x, fma(s, a1y, a0x));
                double ti = fma(c, a1y, fma(-s, a1x, a0y));
                if (hi && (((uint32_t)__popcll(gfull & d.nbr_mask[j]) & 1u) != 0)) {
                    tr = -tr;
                    ti = -ti;
                }
                v.x = tr;
                v.y = ti;
            }
        }
        if (((uint32_t)i & dead_mask) == 0) state[i] = v;
    }
}

// Measurement of a SHARD slot (the pair partner lives on another GPU), fused with the transfer:
// the kernel reads the partner's half straight from peer memory over NVLink and leaves both
// results local, which moves the appended qubit into the top LOCAL slot and the qubit that
// lived there into the shard slot (the host swaps the two slot labels afterwards).
//   role 0 (this rank has shard bit 0, keeps its low half H0):
//       t = own[i] + e^{-i th} peer[i];  own[i] = t;  spare[i] = +-t   (spare becomes the new H1)
//   role 1 (shard bit 1, keeps its high half H1):
//       t = peer[i] + e^{-i th} own[i];  spare[i] = t (the new H0);  own[i] = +-t
//   role 2 (tail, nothing appended; this rank has shard bit 0 and survives): whole shard,
//       own[i] = own[i] + e^{-i th} peer[i]
struct ExchangeParams {
    double2* own;
    const double2* peer;
    double2* spare;
    int32_t role;
    double c, s, scale;
    uint64_t nbr_mask;  // over the bits of i
    uint32_t const_parity;
    uint64_t n;         // live elements of the half (dead slots squeezed out of the loop index)
    int32_t n_ranges;
    uint32_t range_pos[MBQC_STREAM_MAX_RANGES];
    uint32_t range_width[MBQC_STREAM_MAX_RANGES];
};

__global__ void __launch_bounds__(256) stream_exchange_kernel(const __grid_constant__ ExchangeParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.n; e += stride) {
        uint64_t i = e;
        for (int r = 0; r < p.n_ranges; ++r) i = insert_zero_field(i, p.range_pos[r], p.range_width[r]);
        const double2 mine = p.own[i];
        const double2 theirs = p.peer[i];  // NVLink peer load
        const double2 a0 = (p.role == 1) ? theirs : mine;  // bit-0 partner
        const double2 a1 = (p.role == 1) ? mine : theirs;  // bit-1 partner
        double2 t;
        t.x = fma(p.c, a1.x, fma(p.s, a1.y, a0.x)) * p.scale;
        t.y = fma(p.c, a1.y, fma(-p.s, a1.x, a0.y)) * p.scale;
        if (p.role == 2) {
            p.own[i] = t;
            continue;
        }
        const uint32_t sg = (p.const_parity ^ ((uint32_t)__popcll(i & p.nbr_mask) & 1u)) << 31;
        const double2 ts = make_double2(flip_sign(t.x, sg), flip_sign(t.y, sg));
        if (p.role == 0) {
            p.own[i] = t;
            p.spare[i] = ts;
        } else {
            p.spare[i] = t;
            p.own[i] = ts;
        }
    }
}

struct StreamInitParams {
    double2* state;
    uint64_t n_local;
    uint64_t index_or;
    SeedDev seed;
};

__global__ void __launch_bounds__(256) stream_init_kernel(const __grid_constant__ StreamInitParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_local; i += stride)
        p.state[i] = seed_amplitude(p.seed, i | p.index_or);
}

struct StreamGatherParams {
    const double2* state;
    double2* out;
    uint64_t index_or, local_mask;  // entries whose high bits equal index_or belong to this rank
    int32_t n_out;
    int32_t out_slot[kMaxIO];
};

__global__ void stream_gather_kernel(const __grid_constant__ StreamGatherParams p) {
    const uint32_t no = 1u << p.n_out;
    for (uint32_t o = blockIdx.x * blockDim.x + threadIdx.x; o < no; o += gridDim.x * blockDim.x) {
        uint64_t idx = 0;
        for (int q = 0; q < p.n_out; ++q) idx |= (uint64_t)((o >> (p.n_out - 1 - q)) & 1u) << p.out_slot[q];
        if ((idx & ~p.local_mask) == p.index_or) p.out[o] = p.state[idx & p.local_mask];
    }
}

}  // namespace mbqc
