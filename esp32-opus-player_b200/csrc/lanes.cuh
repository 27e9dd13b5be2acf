// lanes.cuh -- the "one warp works on one frame" execution group.
//
// Device: 32 lanes of a warp, exchanging data through shared memory and warp reductions.
// Host (tests/hostsim only): a single lane, so the very same stage bodies run serially and can be
// checked against the reference oracle on a CPU-only box.
#pragma once
#include "fixmath.cuh"

namespace opusb200 {

#if defined(__CUDACC__)
enum { kLanes = 32 };
struct Lanes {
    int lane;
    __device__ __forceinline__ Lanes() : lane((int)(threadIdx.x & 31)) {}
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ s32 sum(s32 v) const { return (s32)__reduce_add_sync(0xffffffffu, (unsigned)v); }
    __device__ __forceinline__ u32 any_or(u32 v) const { return __reduce_or_sync(0xffffffffu, v); }
    __device__ __forceinline__ s32 bcast(s32 v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    template <class T>
    __device__ __forceinline__ T bcast_op(const T& v, int src) const {  // 16-byte POD broadcast
        static_assert(sizeof(T) == 16, "bcast_op: 16-byte records");
        uint4 u = *reinterpret_cast<const uint4*>(&v);
        u.x = __shfl_sync(0xffffffffu, u.x, src); u.y = __shfl_sync(0xffffffffu, u.y, src);
        u.z = __shfl_sync(0xffffffffu, u.z, src); u.w = __shfl_sync(0xffffffffu, u.w, src);
        T r;
        *reinterpret_cast<uint4*>(&r) = u;
        return r;
    }
};
#else
enum { kLanes = 1 };
struct Lanes {
    int lane;
    Lanes() : lane(0) {}
    void sync() const {}
    s32 sum(s32 v) const { return v; }
    u32 any_or(u32 v) const { return v; }
    s32 bcast(s32 v, int) const { return v; }
    template <class T>
    T bcast_op(const T& v, int) const { return v; }
};
#endif

}  // namespace opusb200
