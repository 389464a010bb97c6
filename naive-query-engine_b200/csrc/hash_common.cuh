// hash_common.cuh -- device hashing and f64 <-> order-preserving u64 keys
#pragma once
#include <stdint.h>

// The reference hashes join keys with XxHash64 and std's SipHash (hash_join.rs:68-70),
// but only real-value equality is observable in the results, so the device tables use
// a cheaper 64-bit finaliser (splitmix64 / murmur3 fmix style).
__host__ __device__ __forceinline__ uint64_t nqe_mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// splitmix64 finaliser used by the synthetic data generator
__host__ __device__ __forceinline__ uint64_t nqe_splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

// OrderedFloat<f64> total order (ordered-float 3.0.0, aggregate/max.rs:30,48-51):
// -inf < ... < -0.0 == 0.0 < ... < +inf < NaN, all NaNs equal.  Encoded as a u64
// whose unsigned order is that order (every NaN maps to the maximum key), so that
// max/min become atomicMax/atomicMin.  -0.0 orders just below +0.0 here; the
// reference keeps whichever zero came first (a sign-of-zero-only difference).
constexpr uint64_t NQE_ORD_NAN = 0xFFFFFFFFFFFFFFFFULL;
__host__ __device__ __forceinline__ uint64_t nqe_f64_to_ord(double d) {
    if (d != d) return NQE_ORD_NAN;
#ifdef __CUDA_ARCH__
    uint64_t b = (uint64_t)__double_as_longlong(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__host__ __device__ __forceinline__ double nqe_ord_to_f64(uint64_t k) {
    uint64_t b;
    if (k == NQE_ORD_NAN) b = 0x7FF8000000000000ULL;
    else b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
