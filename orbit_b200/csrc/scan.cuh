// scan.cuh — device-wide ordered compaction support: single-pass scan with decoupled look-back.
//
// Replaces the reference's `atomicAdd` appends (entity_cull.comp:211, meshlet_cull.comp:228,
// active_cluster_compaction.comp:30, light_culling.comp:133) with a deterministic exclusive prefix over tiles.
// A tile publishes {epoch, flag, value} as ONE 64-bit word, so no fence is needed between flag and value, and
// the epoch makes every descriptor of an earlier launch read as "not ready" — no memset between launches.
// The epoch lives in device memory and is advanced by the last CTA of every launch (graph-replay safe).
// Tiles are handed out by an atomic ticket, so a tile's predecessors are always owned by running CTAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace orbit {

// Programmatic dependent launch (PDL): every kernel of the pipeline waits for its predecessor's results before its
// first global access (`wait`) and lets its successor be scheduled once its own main work is done
// (`launch_dependents` near the end), so the launch ramp of each of the 7 small dependent kernels of a frame
// overlaps the tail of the previous one. Both are no-ops when a kernel is launched without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Development aid (never in the shipped library): -DORBIT_TRACE makes thread 0 of every CTA write %globaltimer stamps at
// named points of a kernel into a buffer owned by the context (orbit_debug_trace): every LAUNCH gets its own block of
// 1024 CTAs x 16 slots (16 blocks, handed out round robin by the host); slot 15 of CTA 0 holds the kernel id.
// tools/trace_frame.py turns the blocks into a timeline of the frame. Without the flag the macro expands to nothing.
#ifdef ORBIT_TRACE
__device__ __forceinline__ unsigned long long trace_time() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define ORBIT_TRACE_STAMP(trace, kernel_id, slot) \
    do { if ((trace) != nullptr && threadIdx.x == 0) { \
        (trace)[(size_t)min(blockIdx.x, 1023u) * 16u + (slot)] = orbit::trace_time(); \
        if (blockIdx.x == 0) (trace)[15] = 1000ull + (kernel_id); } } while (0)
// stamp taken only once the 32-bit value `dep` has arrived (the asm reads it, so the warp stalls on its scoreboard first)
__device__ __forceinline__ unsigned long long trace_time_after(unsigned int dep) {
    unsigned long long t; asm volatile("{ .reg .b32 tmp; mov.b32 tmp, %1; mov.u64 %0, %%globaltimer; }" : "=l"(t) : "r"(dep)); return t; }
#define ORBIT_TRACE_STAMP_AFTER(trace, kernel_id, slot, dep) \
    do { if ((trace) != nullptr && threadIdx.x == 0) (trace)[(size_t)min(blockIdx.x, 1023u) * 16u + (slot)] = orbit::trace_time_after((unsigned int)(dep)); } while (0)
#define ORBIT_TRACE_VALUE(trace, slot, value) \
    do { if ((trace) != nullptr && threadIdx.x == 0) (trace)[(size_t)min(blockIdx.x, 1023u) * 16u + (slot)] = (unsigned long long)(value); } while (0)
#else
#define ORBIT_TRACE_STAMP(trace, kernel_id, slot) do { } while (0)
#define ORBIT_TRACE_STAMP_AFTER(trace, kernel_id, slot, dep) do { } while (0)
#define ORBIT_TRACE_VALUE(trace, slot, value) do { } while (0)
#endif

struct ScanState {
    unsigned long long* trace;   // development timeline buffer (nullptr unless built with -DORBIT_TRACE)
    unsigned long long* status;  // one descriptor per tile
    unsigned int* ticket;        // next tile to hand out
    unsigned int* done;          // CTAs that finished (the last one resets ticket/done for the next launch)
    unsigned int* epoch;         // launch id in DEVICE memory (1..2^30-1), advanced by the last CTA of each launch,
                                 // so a CUDA graph that replays the same kernel node still gets a fresh epoch
};

// Every thread of a scan kernel reads the epoch once at kernel start (it only changes at the very end of a launch).
__device__ __forceinline__ unsigned int scan_epoch(const ScanState& st) { return __ldcg(st.epoch); }

enum : unsigned int { kFlagAggregate = 1u, kFlagInclusive = 2u };

__device__ __forceinline__ unsigned long long pack_status(unsigned int epoch, unsigned int flag, unsigned int value) {
    return ((unsigned long long)((epoch << 2) | flag) << 32) | value;
}
__device__ __forceinline__ void publish(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peek(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by all 32 lanes of ONE warp of the CTA that owns `tile`. Returns the exclusive prefix of `aggregate`
// over tiles 0..tile-1 (same value in every lane) and publishes this tile's inclusive prefix.
__device__ __forceinline__ unsigned int lookback_exclusive(const ScanState& st, unsigned int epoch, unsigned int tile, unsigned int aggregate) {
    const unsigned int lane = threadIdx.x & 31u;
    if (tile == 0u) {
        if (lane == 0u) publish(st.status, pack_status(epoch, kFlagInclusive, aggregate));
        return 0u;
    }
    if (lane == 0u) publish(st.status + tile, pack_status(epoch, kFlagAggregate, aggregate));
    unsigned int exclusive = 0u;
    int pos = (int)tile - 1;
    while (true) {
        const int idx = pos - (int)lane;
        unsigned int flag, value;
        if (idx >= 0) {
            while (true) {
                unsigned long long w = peek(st.status + idx);
                unsigned int hi = (unsigned int)(w >> 32);
                if ((hi >> 2) == epoch && (hi & 3u) != 0u) { flag = hi & 3u; value = (unsigned int)w; break; }
                __nanosleep(32);
            }
        } else {
            flag = kFlagInclusive; value = 0u;  // virtual tile -1: inclusive prefix 0
        }
        const unsigned int incl = __ballot_sync(0xFFFFFFFFu, flag == kFlagInclusive);
        // lanes are ordered nearest predecessor first; stop at the first inclusive descriptor
        const unsigned int first = (unsigned int)__ffs((int)incl) - 1u;  // incl != 0 eventually (virtual tile)
        const bool take = (incl == 0u) || (lane <= first);
        unsigned int v = take ? value : 0u;
        v = __reduce_add_sync(0xFFFFFFFFu, v);
        exclusive += v;
        if (incl != 0u) break;
        pos -= 32;
    }
    if (lane == 0u) publish(st.status + tile, pack_status(epoch, kFlagInclusive, exclusive + aggregate));
    return exclusive;
}

// Flat gather: sum of the aggregates published (flag kFlagAggregate, this epoch) by CTAs [0, n), computed by ONE warp.
// All loads of a batch are issued before any is checked, so a fully published prefix costs one L2 round trip per
// 256 CTAs instead of one per CTA; descriptors that are not ready yet are re-polled.
__device__ __forceinline__ unsigned int gather_lower_aggregates(const ScanState& st, unsigned int epoch, unsigned int n) {
    const unsigned int lane = threadIdx.x & 31u;
    unsigned int sum = 0u;
    for (unsigned int base = 0u; base < n; base += 256u) {
        unsigned long long w[8];
        unsigned int pending = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned int i = base + (unsigned int)k * 32u + lane;
            w[k] = 0ull;
            if (i < n) { w[k] = peek(st.status + i); pending |= 1u << k; }
        }
#ifdef ORBIT_TRACE
        unsigned int trace_iters = 0u;
#endif
        while (pending) {
#ifdef ORBIT_TRACE
            if (trace_iters == 0u) ORBIT_TRACE_STAMP(st.trace, 0, 9);
            ++trace_iters;
            ORBIT_TRACE_VALUE(st.trace, 8, trace_iters);
#endif
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if ((pending >> k) & 1u) {
                    if ((unsigned int)(w[k] >> 34) == epoch) { sum += (unsigned int)w[k]; pending &= ~(1u << k); }
                    else w[k] = peek(st.status + base + (unsigned int)k * 32u + lane);
                }
            }
        }
    }
    return __reduce_add_sync(0xFFFFFFFFu, sum);
}

// Every CTA calls this once on exit (one thread). The last CTA re-arms the ticket for the next launch.
// No fence is needed: the control words touched here do not depend on the visibility of the kernel's data writes
// (the kernel boundary publishes those), and a fence would add a full store-drain round trip to every CTA's exit.
__device__ __forceinline__ void scan_cta_exit(const ScanState& st, unsigned int epoch, unsigned int* also_zero = nullptr) {
    unsigned int prev = atomicAdd(st.done, 1u);
    if (prev + 1u == gridDim.x) {
        *st.ticket = 0u;
        *st.done = 0u;
        if (also_zero) *also_zero = 0u;
        unsigned int next = (epoch + 1u) & 0x3FFFFFFFu;
        *st.epoch = next ? next : 1u;
    }
}

}  // namespace orbit
