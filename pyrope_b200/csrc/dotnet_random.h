// dotnet_random.h — System.Random(int seed) as the reference's k-means initialisation needs it
// (KMeansUtils.cs:16-20: data.OrderBy(_ => rnd.Next()).Take(k)).  The generator is the .NET BCL's
// seeded compatibility PRNG (Knuth subtractive, Net5CompatSeedImpl) — a runtime dependency of the
// reference, not part of its source tree; restated from the published algorithm.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace pyrope {

class DotNetRandom {
public:
    explicit DotNetRandom(int32_t seed) {
        const int32_t kMax = 2147483647;
        int32_t sub = (seed == INT32_MIN) ? kMax : (seed < 0 ? -seed : seed);
        int32_t mj = 161803398 - sub;
        int32_t mk = 1;
        for (int i = 0; i < 56; ++i) sa_[i] = 0;
        sa_[55] = mj;
        int ii = 0;
        for (int i = 1; i < 55; ++i) {
            if ((ii += 21) >= 55) ii -= 55;
            sa_[ii] = mk;
            mk = mj - mk;
            if (mk < 0) mk += kMax;
            mj = sa_[ii];
        }
        for (int k = 1; k < 5; ++k)
            for (int i = 1; i < 56; ++i) {
                int n = i + 30;
                if (n >= 55) n -= 55;
                sa_[i] = (int32_t)((uint32_t)sa_[i] - (uint32_t)sa_[1 + n]);
                if (sa_[i] < 0) sa_[i] += kMax;
            }
        inext_ = 0;
        inextp_ = 21;
    }
    int32_t Next() {
        const int32_t kMax = 2147483647;
        if (++inext_ >= 56) inext_ = 1;
        if (++inextp_ >= 56) inextp_ = 1;
        int32_t r = (int32_t)((uint32_t)sa_[inext_] - (uint32_t)sa_[inextp_]);
        if (r == kMax) r--;
        if (r < 0) r += kMax;
        sa_[inext_] = r;
        return r;
    }

private:
    int32_t sa_[56];
    int inext_, inextp_;
};

// Indices of the first k elements of data.OrderBy(_ => rnd.Next()): one key per element drawn in
// order, stable ascending sort (LINQ OrderBy is stable), first k.
inline std::vector<int64_t> kmeans_init_indices(int64_t n, int k, int32_t seed) {
    DotNetRandom rnd(seed);
    std::vector<std::pair<int32_t, int64_t>> keys((size_t)n);
    for (int64_t i = 0; i < n; ++i) keys[(size_t)i] = {rnd.Next(), i};
    if ((int64_t)k < n) {
        std::nth_element(keys.begin(), keys.begin() + k, keys.end());
        std::sort(keys.begin(), keys.begin() + k);
    } else {
        std::sort(keys.begin(), keys.end());
    }
    std::vector<int64_t> out((size_t)k);
    for (int i = 0; i < k; ++i) out[(size_t)i] = keys[(size_t)i].second;
    return out;
}

}  // namespace pyrope
