// Framebuffer exchange over peer memory (one process per GPU on an NVLink / NVSwitch box).
//
// Instead of an NCCL gather into a staging buffer followed by an assembly pass, the root's row-major frame is mapped
// into every rank (CUDA IPC) and each rank stores the tiles it owns straight into it: the scatter kernel below writes
// whole tile rows (256 B per 64-pixel row) through NVLink and its last CTA bumps an arrival counter in the root's
// memory.  The root's stream waits for world-1 arrivals per frame (a one-thread kernel polling its own memory) and, once
// the frame has been consumed, publishes a credit the peers wait for before overwriting the buffer.  No SM of the root
// moves any pixel of the other ranks, and nothing synchronises with the host.
//
//   exchange block (root's device memory, or pinned host memory shared by all processes):
//       [ frame u32[W*H] | pad | credit u32 | pad | arrived u32[64] ]
//   lane sequence numbers start at 1:  rank r stores arrived[r] = seq once its tiles of frame seq are in place (a plain store
//   behind a system-scope fence: idempotent, needs no atomics, so the block may as well live in host memory), the root waits
//   for arrived[1 .. world-1] >= seq, publishes credit = seq when the frame has been consumed; peers wait credit >= seq-1.
//
// With the block in pinned, mapped HOST memory shared by all ranks (hdt_exchange_attach_host) the same kernels make every
// rank push its own tiles over its own PCIe link: the assembled frame materialises in host memory without passing through
// rank 0's GPU or its single link.
#pragma once
#include "hdt_device.cuh"

namespace hdt {

// What a fused final pass needs: rank 0's frame (nullptr = not fused), this context's last-CTA counter, and rank 0's
// arrival counter (nullptr on rank 0 itself, which does not signal).
struct ExchangeOut { u32* frame; u32* ctasDone; u32* arrived; u32 seq; };

constexpr u32 kMaxExchangeRanks = 64;
struct ExchangeCounters { u32 credit; u32 pad[63]; u32 arrived[kMaxExchangeRanks]; };   // 512 B, lives behind the frame

__device__ __forceinline__ u32 load_volatile(const u32* p) { return *reinterpret_cast<const volatile u32*>(p); }

// One thread: wait until words[i] - target >= 0 (wrap-safe) for every i < nWords.  After maxCycles SM cycles (0 = wait for ever; default
// ~20 s, HDT_OPT_EXCHANGE_TIMEOUT_MS) it gives up: *timedOut (host-mapped) is raised, so a lost peer shows up as an error
// at the next host-synchronising call instead of a hung box, and *abortFlag (device) makes the kernels queued behind it
// skip their stores and signals -- a frame that timed out is dropped, never half-written over one still being read.
__global__ void exchange_wait_kernel(const u32* words, u32 nWords, u32 target, unsigned long long maxCycles, u32* timedOut, u32* abortFlag)
{
    const long long t0 = clock64();
    for (u32 i = 0; i < nWords; ++i) {
        while (int(load_volatile(words + i) - target) < 0) {
            __nanosleep(100);
            if (maxCycles && (unsigned long long)(clock64() - t0) > maxCycles) {
                *reinterpret_cast<volatile u32*>(timedOut) = 1;
                *reinterpret_cast<volatile u32*>(abortFlag) = 1;
                i = nWords;
                break;
            }
        }
    }
    __threadfence_system();
}

__global__ void exchange_publish_kernel(u32* counter, u32 value)
{
    __threadfence_system();
    *reinterpret_cast<volatile u32*>(counter) = value;
}

__global__ void exchange_signal_kernel(u32* arrived, u32 seq, const u32* abortFlag)   // a rank that owns no tile still has to arrive
{
    if (load_volatile(abortFlag)) return;
    __threadfence_system();
    *reinterpret_cast<volatile u32*>(arrived) = seq;
}

// End of a kernel that stored into the shared frame: make the CTA's stores visible system-wide, count the CTA, and let the
// last one publish arrived[rank] = seq.  Every thread of the CTA must call it.
__device__ __forceinline__ void exchange_signal_last_cta(u32* __restrict__ ctasDone, u32* arrived, u32 seq)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const u32 prev = atomicAdd(ctasDone, 1u);
        if (prev == gridDim.x - 1) {
            *ctasDone = 0;                            // ready for the next launch (stream order separates launches)
            __threadfence_system();
            if (arrived) *reinterpret_cast<volatile u32*>(arrived) = seq;
        }
    }
}

// This rank's compact tiles (each (1<<tileLog2)^2 pixels, row-major, owned tiles back to back) -> the row-major frame,
// which may live in another GPU's memory.  One CTA per 16 tile rows; the last CTA to finish signals `arrivals`.
__global__ void __launch_bounds__(256) exchange_scatter_kernel(const u32* __restrict__ compact, u32* __restrict__ frame, const PixelMap map,
                                                                u32* __restrict__ ctasDone, u32* arrived, u32 seq, const u32* abortFlag)
{
    if (load_volatile(abortFlag)) return;             // the wait in front of this launch gave up: drop the frame
    const u32 T = 1u << map.tileLog2, rowsPerCta = 16, ctasPerTile = T / rowsPerCta;
    const u32 slot = blockIdx.x / ctasPerTile, row0 = (blockIdx.x % ctasPerTile) * rowsPerCta;
    const u32 t = map.rank + slot * map.world;
    const u32 x0 = (t % map.tilesX) * T, y0 = (t / map.tilesX) * T;
    const u32* src = compact + (u64(slot) << (2 * map.tileLog2));
    if ((map.width & 3) == 0) {
        const u32 vecPerRow = T / 4;
        for (u32 i = threadIdx.x; i < rowsPerCta * vecPerRow; i += blockDim.x) {
            const u32 r = row0 + i / vecPerRow, c = (i % vecPerRow) * 4;
            const u32 x = x0 + c, y = y0 + r;
            if (y < map.height && x < map.width)      // width % 4 == 0: a vector is inside or outside as a whole
                *reinterpret_cast<uint4*>(frame + u64(y) * map.width + x) = *reinterpret_cast<const uint4*>(src + r * T + c);
        }
    } else {
        for (u32 i = threadIdx.x; i < rowsPerCta * T; i += blockDim.x) {
            const u32 r = row0 + i / T, c = i % T;
            const u32 x = x0 + c, y = y0 + r;
            if (y < map.height && x < map.width) frame[u64(y) * map.width + x] = src[r * T + c];
        }
    }
    exchange_signal_last_cta(ctasDone, arrived, seq);
}

}  // namespace hdt
