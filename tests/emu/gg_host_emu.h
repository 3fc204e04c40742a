/*
 * gg_host_emu.h — TEST INFRASTRUCTURE.  Host stand-ins that let greengage_b200/csrc/gg_device.cuh (compiled with
 * -DGG_HOST_EMU by g++) run on a CPU: "shared memory" is a byte array addressed by the same 32-bit offsets the device code
 * uses, the round-to-nearest double intrinsics are the plain IEEE operations (build with -ffp-contract=off), __ldg is a load.
 * Everything else of the tuple walk and the accumulator-machine interpreter is the product's own source, unchanged.
 */
#ifndef GG_HOST_EMU_H
#define GG_HOST_EMU_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __noinline__ __attribute__((noinline))

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long) (((unsigned __int128) a * b) >> 64); }
static inline long long __mul64hi(long long a, long long b) { return (long long) (((__int128) a * b) >> 64); }

#define GG_EMU_SMEM_BYTES (256 * 1024)
extern uint8_t gg_emu_smem[GG_EMU_SMEM_BYTES];

namespace ggd {
static inline uint32_t lds8(uint32_t a) { return gg_emu_smem[a]; }
static inline uint32_t lds16(uint32_t a) { uint16_t v; memcpy(&v, gg_emu_smem + a, 2); return v; }
static inline uint32_t lds32(uint32_t a) { uint32_t v; memcpy(&v, gg_emu_smem + a, 4); return v; }
static inline uint64_t lds64(uint32_t a) { uint64_t v; memcpy(&v, gg_emu_smem + a, 8); return v; }
static inline double ldsf64(uint32_t a) { double v; memcpy(&v, gg_emu_smem + a, 8); return v; }
static inline void sts16(uint32_t a, uint32_t v) { uint16_t x = (uint16_t) v; memcpy(gg_emu_smem + a, &x, 2); }
static inline void sts32(uint32_t a, uint32_t v) { memcpy(gg_emu_smem + a, &v, 4); }
static inline void sts64(uint32_t a, uint64_t v) { memcpy(gg_emu_smem + a, &v, 8); }
static inline void stsf64(uint32_t a, double v) { memcpy(gg_emu_smem + a, &v, 8); }
}
#endif
