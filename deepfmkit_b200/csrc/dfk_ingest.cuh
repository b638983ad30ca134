// Raw-data ingest on the device.
//   * DFMSWPM text records (core.py:259-286: pd.read_csv(sep=' ', skiprows=13, usecols=[c])): the file's bytes are
//     parsed where they land -- rows located by a newline scan, every field converted by a restatement of the float
//     converter the reference's pandas call uses (pandas' C parser default, precise_xstrtod in
//     pandas/_libs/src/parser/tokenizer.c; pinned bit-for-bit against pandas 3.0.2 by tests/golden/ingest_text.npz).
//   * binary acquisition formats (int16 / int32 / float32 / float64, time-major interleaved or channel-major): widened
//     to the fp64 channel-major record the fitters read.
#pragma once
#include "dfk_common.cuh"

namespace dfk {

constexpr int kTxtThreads = 256;
constexpr int kTxtChunk = 64 * 1024;  // bytes per counting chunk: 16 iterations of 256 threads x 16 B

DFK_D bool txt_is_blank(unsigned char c) { return c == ' ' || c == '\t' || c == '\r'; }

// Does a data row begin at byte i?  (i is 0 or follows a '\n'.)  Lines holding only blanks are skipped, as
// read_csv's skip_blank_lines does for empty ones.
DFK_D bool txt_row_begins(const unsigned char* __restrict__ s, long long n, long long i) {
    while (i < n && txt_is_blank(s[i])) ++i;
    return i < n && s[i] != '\n';
}

// number of row starts announced by the 16 bytes at [i, i+16): a '\n' at byte j announces a row at j+1
DFK_D int txt_count16(const unsigned char* __restrict__ s, long long n, long long i) {
    int cnt = 0;
    if (i + 16 <= n && (reinterpret_cast<uintptr_t>(s + i) & 15u) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(s + i);
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (((w[k] >> (8 * b)) & 0xffu) == '\n' && txt_row_begins(s, n, i + 4 * k + b + 1)) ++cnt;
        }
    } else {
        for (int b = 0; b < 16 && i + b < n; ++b)
            if (s[i + b] == '\n' && txt_row_begins(s, n, i + b + 1)) ++cnt;
    }
    return cnt;
}

// Pass 1: rows announced inside each chunk.
__global__ void __launch_bounds__(kTxtThreads) txt_count_kernel(const unsigned char* __restrict__ s, long long n,
                                                               long long nchunks, unsigned* __restrict__ counts) {
    __shared__ int part[kTxtThreads / 32];
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long base = ch * kTxtChunk;
        int cnt = 0;
        for (int it = 0; it < kTxtChunk / (kTxtThreads * 16); ++it) {
            const long long i = base + (static_cast<long long>(it) * kTxtThreads + threadIdx.x) * 16;
            if (i < n) cnt += txt_count16(s, n, i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
#pragma unroll
            for (int w = 0; w < kTxtThreads / 32; ++w) t += part[w];
            counts[ch] = static_cast<unsigned>(t);
        }
        __syncthreads();
    }
}

// Exclusive scan of the chunk counts (one CTA; a 70 GB file has ~1e6 chunks).  offsets[nchunks] = total.
// `first` = 1 when a row begins at byte 0 (it is announced by no newline).
__global__ void __launch_bounds__(1024) txt_scan_kernel(const unsigned* __restrict__ counts, long long nchunks,
                                                       long long first, long long* __restrict__ offsets) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = first;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long base = 0; base < nchunks; base += 1024) {
        const long long i = base + threadIdx.x;
        const long long v = i < nchunks ? counts[i] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_tot[warp - 1] : 0) + inc - v;
        if (i < nchunks) offsets[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[nchunks] = carry;
}

// Pass 2: byte offset of every row, in file order.
__global__ void __launch_bounds__(kTxtThreads) txt_starts_kernel(const unsigned char* __restrict__ s, long long n,
                                                                long long nchunks, const long long* __restrict__ offsets,
                                                                long long first, long long* __restrict__ starts) {
    __shared__ int warp_tot[kTxtThreads / 32];
    __shared__ int iter_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && first) starts[0] = 0;
    for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        const long long base = ch * kTxtChunk;
        long long rank0 = offsets[ch];
        for (int it = 0; it < kTxtChunk / (kTxtThreads * 16); ++it) {
            const long long i = base + (static_cast<long long>(it) * kTxtThreads + threadIdx.x) * 16;
            const int cnt = i < n ? txt_count16(s, n, i) : 0;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) warp_tot[warp] = inc;
            __syncthreads();
            int before = inc - cnt;
            for (int w = 0; w < warp; ++w) before += warp_tot[w];
            if (threadIdx.x == kTxtThreads - 1) iter_tot = before + cnt;
            if (cnt) {
                long long r = rank0 + before;
                for (int b = 0; b < 16 && i + b < n; ++b)
                    if (s[i + b] == '\n' && txt_row_begins(s, n, i + b + 1)) starts[r++] = i + b + 1;
            }
            __syncthreads();
            rank0 += iter_tot;
            __syncthreads();
        }
    }
}

// ---- the float converter --------------------------------------------------------------------------------------
// precise_xstrtod: up to 17 significant digits accumulated as number = number * 10 + digit in fp64 (further integer
// digits only bump the exponent, further decimals are dropped), then ONE multiplication or division by a tabulated
// power of ten.  Not correctly rounded -- 24 % of repr() strings come back 1 ulp off -- and reproduced as such.
// The first 15 digits are exact in either arithmetic, so they are gathered in an integer.
struct TxtNumber {
    double value;
    int ok;        // a number was read
    long long end;  // first byte after it
};

DFK_D TxtNumber txt_strtod(const unsigned char* __restrict__ s, long long p, long long n, const double* __restrict__ pow10) {
    TxtNumber out;
    out.ok = 0;
    bool neg = false;
    if (p < n && (s[p] == '-' || s[p] == '+')) {
        neg = s[p] == '-';
        ++p;
    }
    unsigned long long head = 0;  // first 15 digits
    double number = 0.0;
    int exponent = 0, num_digits = 0, num_decimals = 0;
    auto push = [&](int d) {
        if (num_digits < 15) {
            head = head * 10ull + static_cast<unsigned long long>(d);
            if (num_digits == 14) number = static_cast<double>(head);
        } else {
            number = __dadd_rn(__dmul_rn(number, 10.0), static_cast<double>(d));
        }
        ++num_digits;
    };
    while (p < n && s[p] >= '0' && s[p] <= '9') {
        if (num_digits < 17) push(s[p] - '0'); else ++exponent;
        ++p;
    }
    if (p < n && s[p] == '.') {
        ++p;
        while (num_digits < 17 && p < n && s[p] >= '0' && s[p] <= '9') {
            push(s[p] - '0');
            ++p;
            ++num_decimals;
        }
        if (num_digits >= 17)
            while (p < n && s[p] >= '0' && s[p] <= '9') ++p;
        exponent -= num_decimals;
    }
    out.end = p;
    if (num_digits == 0) {
        out.value = __longlong_as_double(0x7ff8000000000000ll);
        return out;
    }
    if (num_digits < 15) number = static_cast<double>(head);
    if (neg) number = -number;
    if (p < n && (s[p] == 'e' || s[p] == 'E')) {
        long long q = p + 1;
        bool eneg = false;
        if (q < n && (s[q] == '-' || s[q] == '+')) {
            eneg = s[q] == '-';
            ++q;
        }
        int e = 0, nd = 0;
        while (q < n && s[q] >= '0' && s[q] <= '9') {
            if (e < 100000) e = e * 10 + (s[q] - '0');
            ++nd;
            ++q;
        }
        if (nd) {
            exponent += eneg ? -e : e;
            p = q;
        }
    }
    out.end = p;
    out.ok = 1;
    if (exponent > 308) {
        number = neg ? -__longlong_as_double(0x7ff0000000000000ll) : __longlong_as_double(0x7ff0000000000000ll);
    } else if (exponent > 0) {
        number = __dmul_rn(number, pow10[exponent]);
    } else if (exponent < -308) {
        if (exponent < -616) {
            number = neg ? -0.0 : 0.0;
        } else {
            number = __ddiv_rn(number, pow10[-308 - exponent]);
            number = __ddiv_rn(number, pow10[308]);
        }
    } else {
        number = __ddiv_rn(number, pow10[-exponent]);
    }
    out.value = number;
    return out;
}

struct TxtParse {
    const unsigned char* s;
    long long n;
    const long long* starts;
    long long nrows;
    int ncols;                  // columns wanted
    const int* usecols;         // file column of wanted column c (ascending), or nullptr for 0..ncols-1
    double* out;                // wanted column c, row r at out[c * ld_c + r]
    long long ld_c;
    const double* pow10;        // 1e0 .. 1e308
    unsigned long long* nbad;   // fields that were missing or not numbers (stored as NaN)
};

// Pass 3: one thread per row.
__global__ void __launch_bounds__(kTxtThreads) txt_parse_kernel(const TxtParse P) {
    __shared__ double p10[309];
    for (int i = threadIdx.x; i < 309; i += kTxtThreads) p10[i] = P.pow10[i];
    __syncthreads();
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned bad = 0;
    for (long long r = blockIdx.x * static_cast<long long>(kTxtThreads) + threadIdx.x; r < P.nrows;
         r += static_cast<long long>(gridDim.x) * kTxtThreads) {
        long long p = P.starts[r];
        int want = 0;  // next wanted column
        int col = 0;   // file column of the field at p
        while (want < P.ncols) {
            while (p < P.n && txt_is_blank(P.s[p])) ++p;
            if (p >= P.n || P.s[p] == '\n') break;
            const int target = P.usecols ? P.usecols[want] : want;
            if (col == target) {
                TxtNumber t = txt_strtod(P.s, p, P.n, p10);
                p = t.end;
                // a field must end at a blank or the line's end; anything else is not a number
                if (!t.ok || (p < P.n && !txt_is_blank(P.s[p]) && P.s[p] != '\n')) {
                    t.value = nan;
                    ++bad;
                    while (p < P.n && !txt_is_blank(P.s[p]) && P.s[p] != '\n') ++p;
                }
                P.out[want * P.ld_c + r] = t.value;
                ++want;
            } else {
                while (p < P.n && !txt_is_blank(P.s[p]) && P.s[p] != '\n') ++p;
            }
            ++col;
        }
        for (; want < P.ncols; ++want) {  // short row
            P.out[want * P.ld_c + r] = nan;
            ++bad;
        }
    }
    if (bad) atomicAdd(P.nbad, static_cast<unsigned long long>(bad));
}

// ---- binary formats ---------------------------------------------------------------------------------------------
// out[c * ld_c + t] = scale * src(t, c) + offset, src time-major interleaved (t * C + c) or channel-major
// (c * T + t).  The interleaved case goes through a shared-memory tile so that both sides stay coalesced.
enum : int { kRawI16 = 0, kRawI32 = 1, kRawF32 = 2, kRawF64 = 3 };

template <class S>
DFK_D double raw_load(const void* __restrict__ src, long long i) { return static_cast<double>(static_cast<const S*>(src)[i]); }

DFK_D double raw_load_any(const void* __restrict__ src, int dtype, long long i) {
    switch (dtype) {
        case kRawI16: return raw_load<short>(src, i);
        case kRawI32: return raw_load<int>(src, i);
        case kRawF32: return raw_load<float>(src, i);
        default: return raw_load<double>(src, i);
    }
}

__global__ void __launch_bounds__(256) widen_channel_major_kernel(const void* __restrict__ src, int dtype, long long T,
                                                                 long long C, long long src_ld, double scale, double offset,
                                                                 double* __restrict__ out, long long ld_c) {
    const long long total = T * C;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256ll) {
        const long long c = i / T, t = i - c * T;
        out[c * ld_c + t] = fma(scale, raw_load_any(src, dtype, c * src_ld + t), offset);
    }
}

// Few channels (the usual acquisition card: 1..16): a thread takes one time step, reads its C interleaved samples
// (consecutive threads read consecutive groups: coalesced) and writes each to its channel's row (coalesced per row).
constexpr int kWidenFewMax = 16;
__global__ void __launch_bounds__(256) widen_few_channels_kernel(const void* __restrict__ src, int dtype, long long T, int C,
                                                                double scale, double offset, double* __restrict__ out,
                                                                long long ld_c) {
    for (long long t = blockIdx.x * 256ll + threadIdx.x; t < T; t += static_cast<long long>(gridDim.x) * 256ll) {
        for (int c = 0; c < C; ++c) out[c * ld_c + t] = fma(scale, raw_load_any(src, dtype, t * C + c), offset);
    }
}

constexpr int kWidenTile = 32;
__global__ void __launch_bounds__(kWidenTile* 8) widen_time_major_kernel(const void* __restrict__ src, int dtype, long long T,
                                                                        long long C, double scale, double offset,
                                                                        double* __restrict__ out, long long ld_c) {
    __shared__ double tile[kWidenTile][kWidenTile + 1];
    const long long tiles_c = (C + kWidenTile - 1) / kWidenTile;
    const long long tiles_t = (T + kWidenTile - 1) / kWidenTile;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (long long tile_id = blockIdx.x; tile_id < tiles_c * tiles_t; tile_id += gridDim.x) {
        const long long t0 = (tile_id / tiles_c) * kWidenTile, c0 = (tile_id % tiles_c) * kWidenTile;
#pragma unroll
        for (int k = 0; k < kWidenTile; k += 8) {  // rows = time, fast index = channel
            const long long t = t0 + ty + k, c = c0 + tx;
            if (t < T && c < C) tile[ty + k][tx] = fma(scale, raw_load_any(src, dtype, t * C + c), offset);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kWidenTile; k += 8) {  // rows = channel, fast index = time
            const long long c = c0 + ty + k, t = t0 + tx;
            if (t < T && c < C) out[c * ld_c + t] = tile[tx][ty + k];
        }
        __syncthreads();
    }
}

}  // namespace dfk
