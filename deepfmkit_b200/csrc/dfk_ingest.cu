// Raw-data ingest entries of the C ABI (include/dfk_b200.h, "raw-data ingest"): DFMSWPM text records and binary
// acquisition formats go from a file (or host memory) through pinned staging buffers to the device, where they are
// parsed / widened into the fp64 channel-major record the fitters read.  Replaces DeepFitFramework.parse_header and
// load_raw (core.py:129-174, 259-286).
#include "dfk_host.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "dfk_ingest.cuh"

namespace {

size_t dtype_bytes(int dtype) {
    switch (dtype) {
        case dfk::kRawI16: return 2;
        case dfk::kRawI32: return 4;
        case dfk::kRawF32: return 4;
        case dfk::kRawF64: return 8;
        default: return 0;
    }
}

// pread of [off, off+n) into dst by a few threads (the page cache copies at memcpy speed per thread)
int parallel_pread(int fd, void* dst, int64_t off, size_t n) {
    CopyPool& pool = CopyPool::instance();
    const int nt = static_cast<int>(std::min<size_t>(static_cast<size_t>(std::min(pool.size(), 8)), n / (8u << 20) + 1));
    std::vector<int> rc(nt, 0);
    const size_t chunk = ((n / nt) + 4095) & ~static_cast<size_t>(4095);
    pool.run(nt, [&](int t) {
        size_t lo = std::min(n, chunk * t), hi = std::min(n, chunk * (t + 1));
        while (lo < hi) {
            const ssize_t got = pread(fd, static_cast<char*>(dst) + lo, hi - lo, off + static_cast<int64_t>(lo));
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) {
                rc[t] = -1;
                return;
            }
            lo += static_cast<size_t>(got);
        }
    });
    for (int r : rc)
        if (r) return fail(DFK_ERR_ARG, "short read at offset %lld: %s", (long long)off, std::strerror(errno));
    return DFK_OK;
}

int ensure_stagers(dfk_ctx* ctx) {
    for (int i = 0; i < host_copy_tuning().stagers; ++i) {
        if (!ctx->stager[i]) {
            DFK_CUDA(cudaHostAlloc(&ctx->stager[i], dfk_ctx::kStageBytes, cudaHostAllocDefault));
            DFK_CUDA(cudaEventCreateWithFlags(&ctx->stager_free[i], cudaEventDisableTiming));
        }
    }
    return DFK_OK;
}

// Bytes [src_off, src_off + bytes) of a source -> device memory at dst, on the copy stream.  The source is host
// memory (`host`; pinned memory goes by DMA directly, pageable memory through the stagers) or a file (`fd`, read
// straight into the stagers so that the file's bytes are touched once on the host).
struct ByteSource {
    const char* host = nullptr;
    int fd = -1;
    bool pageable = true;
};

int source_to_device(dfk_ctx* ctx, const ByteSource& src, int64_t src_off, void* dst, size_t bytes) {
    if (src.host) return copy_slab_to_device(ctx, dst, src.host + src_off, bytes, src.pageable);
    int rc = ensure_stagers(ctx);
    if (rc) return rc;
    int k = 0;
    const HostCopyTuning& tune = host_copy_tuning();
    for (size_t off = 0; off < bytes; off += tune.stage_bytes, k = (k + 1) % tune.stagers) {
        const size_t n = std::min(tune.stage_bytes, bytes - off);
        DFK_CUDA(cudaEventSynchronize(ctx->stager_free[k]));
        rc = parallel_pread(src.fd, ctx->stager[k], src_off + static_cast<int64_t>(off), n);
        if (rc) return rc;
        DFK_CUDA(cudaMemcpyAsync(static_cast<char*>(dst) + off, ctx->stager[k], n, cudaMemcpyHostToDevice, ctx->copy_stream));
        DFK_CUDA(cudaEventRecord(ctx->stager_free[k], ctx->copy_stream));
    }
    return DFK_OK;
}

struct FileCloser {
    int fd;
    ~FileCloser() {
        if (fd >= 0) close(fd);
    }
};

const double* pow10_table(dfk_ctx* ctx, cudaStream_t st, int* rc_out) {
    DevBuf& b = ctx->post[5];
    if (!b.ptr) {
        *rc_out = ensure(ctx, b, sizeof(double) * 309);
        if (*rc_out) return nullptr;
        static double table[309];
        for (int i = 0; i < 309; ++i) {  // the doubles nearest to 1e0 .. 1e308, as the C literals of pandas' table
            char lit[16];
            snprintf(lit, sizeof(lit), "1e%d", i);
            table[i] = std::strtod(lit, nullptr);
        }
        if (cudaMemcpyAsync(b.ptr, table, sizeof(table), cudaMemcpyHostToDevice, st) != cudaSuccess) {
            *rc_out = fail(DFK_ERR_CUDA, "upload of the power-of-ten table failed");
            return nullptr;
        }
    }
    *rc_out = DFK_OK;
    return static_cast<const double*>(b.ptr);
}

// rows of the text now resident in post[3]: chunk counts -> exclusive scan; leaves offsets in post[4]
int index_text(dfk_ctx* ctx, int64_t nbytes, int64_t* nrows_out) {
    cudaStream_t st = ctx->stream();
    const int64_t nchunks = (nbytes + dfk::kTxtChunk - 1) / dfk::kTxtChunk;
    int rc = ensure(ctx, ctx->post[4], sizeof(long long) * (nchunks + 1) + sizeof(unsigned) * nchunks + 64);
    if (rc) return rc;
    long long* offsets = static_cast<long long*>(ctx->post[4].ptr);
    unsigned* counts = reinterpret_cast<unsigned*>(offsets + nchunks + 1);
    const unsigned char* s = static_cast<const unsigned char*>(ctx->post[3].ptr);
    // does a row begin at byte 0?  (decided on the host from the first bytes: no newline announces it)
    unsigned char head[256];
    const size_t nh = static_cast<size_t>(std::min<int64_t>(nbytes, sizeof(head)));
    DFK_CUDA(cudaMemcpyAsync(head, s, nh, cudaMemcpyDeviceToHost, st));
    DFK_CUDA(cudaStreamSynchronize(st));
    size_t i = 0;
    while (i < nh && (head[i] == ' ' || head[i] == '\t' || head[i] == '\r')) ++i;
    int first = (i < nh && head[i] != '\n') ? 1 : 0;
    if (i == nh && nbytes > static_cast<int64_t>(nh)) first = 1;  // 256 blanks: treat as a (malformed) row
    const int grid = static_cast<int>(std::min<int64_t>(nchunks, static_cast<int64_t>(ctx->sm_count) * 8));
    dfk::txt_count_kernel<<<grid, dfk::kTxtThreads, 0, st>>>(s, nbytes, nchunks, counts);
    DFK_CUDA(cudaGetLastError());
    dfk::txt_scan_kernel<<<1, 1024, 0, st>>>(counts, nchunks, first, offsets);
    DFK_CUDA(cudaGetLastError());
    ctx->launches += 2;
    long long total = 0;
    DFK_CUDA(cudaMemcpyAsync(&total, offsets + nchunks, sizeof(long long), cudaMemcpyDeviceToHost, st));
    DFK_CUDA(cudaStreamSynchronize(st));
    ctx->text_bytes = nbytes;
    ctx->text_rows = total;
    ctx->text_chunks = nchunks;
    ctx->text_first = first;
    if (nrows_out) *nrows_out = total;
    return DFK_OK;
}

int load_text(dfk_ctx* ctx, const ByteSource& src, int64_t src_off, int64_t nbytes, int64_t* nrows_out) {
    ctx->text_bytes = ctx->text_rows = ctx->text_chunks = 0;
    if (nbytes <= 0) {
        if (nrows_out) *nrows_out = 0;
        return DFK_OK;
    }
    int rc = ensure(ctx, ctx->post[3], static_cast<size_t>(nbytes) + 64);
    if (rc) return rc;
    HostCallGuard hg(ctx);
    rc = source_to_device(ctx, src, src_off, ctx->post[3].ptr, static_cast<size_t>(nbytes));
    if (rc) return rc;
    DFK_CUDA(cudaEventRecord(ctx->copied[0], ctx->copy_stream));
    DFK_CUDA(cudaStreamWaitEvent(ctx->stream(), ctx->copied[0], 0));
    rc = index_text(ctx, nbytes, nrows_out);
    if (rc) return rc;
    hg.done();
    return DFK_OK;
}

// Python's int() / float() on the digits-and-dots string the reference builds from a header line (core.py:149-150)
bool header_int(const std::string& v, int64_t* out) {
    if (v.empty() || v.find('.') != std::string::npos || v.size() > 18) return false;
    *out = std::strtoll(v.c_str(), nullptr, 10);
    return true;
}
bool header_float(const std::string& v, double* out) {
    if (v.empty() || v == "." || std::count(v.begin(), v.end(), '.') > 1) return false;
    *out = std::strtod(v.c_str(), nullptr);
    return true;
}

}  // namespace

extern "C" {

int dfk_raw_parse_header(const char* path, dfk_raw_header* out) {
    if (!path || !out) return fail(DFK_ERR_ARG, "null pointer");
    FILE* f = std::fopen(path, "rb");
    if (!f) return fail(DFK_ERR_ARG, "cannot open %s: %s", path, std::strerror(errno));
    std::vector<std::string> lines;
    std::string cur;
    int64_t pos = 0, data_offset = -1;
    int c;
    while ((c = std::fgetc(f)) != EOF) {
        ++pos;
        if (c == '\n') {
            lines.push_back(cur);
            cur.clear();
            if (lines.size() == 13) {  // read_csv(skiprows=13): 12 '%' lines and the column-name line
                data_offset = pos;
                break;
            }
        } else {
            cur.push_back(static_cast<char>(c));
        }
    }
    std::fclose(f);
    if (lines.size() < 11) return fail(DFK_ERR_ARG, "%s: header has %zu lines, 11 expected", path, lines.size());
    std::string values[4];
    for (int v = 0; v < 4; ++v)  // lines 2..5: keep the characters in '1234567890.' (core.py:149-150)
        for (char ch : lines[2 + v])
            if ((ch >= '0' && ch <= '9') || ch == '.') values[v].push_back(ch);
    int64_t channr = 0, t0 = 0;
    double f_samp = 0.0, f_mod = 0.0;
    if (!header_int(values[0], &channr) || !header_int(values[1], &t0) || !header_float(values[2], &f_samp) ||
        !header_float(values[3], &f_mod))
        return fail(DFK_ERR_ARG, "%s: header lines 2-5 do not hold channels / start time / f_samp / f_mod", path);
    out->channels = static_cast<int32_t>(channr);
    out->t0 = t0;
    out->f_samp = f_samp;
    out->f_mod = f_mod;
    out->data_offset = data_offset < 0 ? pos : data_offset;
    return DFK_OK;
}

int dfk_text_load_file(dfk_ctx* ctx, const char* path, int64_t byte_offset, int64_t* nbytes_out, int64_t* nrows_out) {
    DFK_ENTER(ctx);
    if (!path) return fail(DFK_ERR_ARG, "null path");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(DFK_ERR_ARG, "cannot open %s: %s", path, std::strerror(errno));
    FileCloser closer{fd};
    struct stat sb;
    if (fstat(fd, &sb) != 0) return fail(DFK_ERR_ARG, "cannot stat %s", path);
    const int64_t nbytes = std::max<int64_t>(0, static_cast<int64_t>(sb.st_size) - byte_offset);
    if (nbytes_out) *nbytes_out = nbytes;
    ByteSource src;
    src.fd = fd;
    return load_text(ctx, src, byte_offset, nbytes, nrows_out);
}

int dfk_text_load_host(dfk_ctx* ctx, const char* text_host, int64_t nbytes, int64_t* nrows_out) {
    DFK_ENTER(ctx);
    if (nbytes < 0 || (!text_host && nbytes > 0)) return fail(DFK_ERR_ARG, "bad text buffer");
    ByteSource src;
    src.host = text_host;
    src.pageable = nbytes > 0 ? is_pageable(text_host) : true;
    return load_text(ctx, src, 0, nbytes, nrows_out);
}

int dfk_text_parse_dev(dfk_ctx* ctx, int32_t ncols, const int32_t* usecols, double* out_dev, int64_t ld_c, int64_t* nbad_out) {
    DFK_ENTER(ctx);
    if (ncols < 1 || ncols > 4096) return fail(DFK_ERR_ARG, "column count %d outside 1..4096", ncols);
    if (nbad_out) *nbad_out = 0;
    const int64_t nrows = ctx->text_rows;
    if (nrows == 0) return DFK_OK;
    if (!out_dev) return fail(DFK_ERR_ARG, "null output");
    if (ld_c < nrows) return fail(DFK_ERR_ARG, "column stride %lld shorter than the %lld rows", (long long)ld_c, (long long)nrows);
    for (int c = 0; usecols && c < ncols; ++c)
        if (usecols[c] < 0 || (c && usecols[c] <= usecols[c - 1])) return fail(DFK_ERR_ARG, "usecols must be ascending and >= 0");
    cudaStream_t st = ctx->stream();
    const size_t cols_bytes = (sizeof(int) * static_cast<size_t>(ncols) + 7) & ~static_cast<size_t>(7);
    int rc = ensure(ctx, ctx->misc, sizeof(long long) * nrows + cols_bytes + sizeof(unsigned long long));
    if (rc) return rc;
    const double* p10 = pow10_table(ctx, st, &rc);
    if (rc) return rc;
    long long* starts = static_cast<long long*>(ctx->misc.ptr);
    int* cols_dev = reinterpret_cast<int*>(starts + nrows);
    unsigned long long* nbad_dev = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(cols_dev) + cols_bytes);
    if (usecols) DFK_CUDA(cudaMemcpyAsync(cols_dev, usecols, sizeof(int) * ncols, cudaMemcpyHostToDevice, st));
    DFK_CUDA(cudaMemsetAsync(nbad_dev, 0, sizeof(unsigned long long), st));
    const unsigned char* s = static_cast<const unsigned char*>(ctx->post[3].ptr);
    const long long* offsets = static_cast<const long long*>(ctx->post[4].ptr);
    const int grid_a = static_cast<int>(std::min<int64_t>(ctx->text_chunks, static_cast<int64_t>(ctx->sm_count) * 8));
    dfk::txt_starts_kernel<<<grid_a, dfk::kTxtThreads, 0, st>>>(s, ctx->text_bytes, ctx->text_chunks, offsets, ctx->text_first,
                                                               starts);
    DFK_CUDA(cudaGetLastError());
    dfk::TxtParse P;
    P.s = s;
    P.n = ctx->text_bytes;
    P.starts = starts;
    P.nrows = nrows;
    P.ncols = ncols;
    P.usecols = usecols ? cols_dev : nullptr;
    P.out = out_dev;
    P.ld_c = ld_c;
    P.pow10 = p10;
    P.nbad = nbad_dev;
    const int grid_b = static_cast<int>(std::min<int64_t>((nrows + dfk::kTxtThreads - 1) / dfk::kTxtThreads,
                                                         static_cast<int64_t>(ctx->sm_count) * 8));
    dfk::txt_parse_kernel<<<grid_b, dfk::kTxtThreads, 0, st>>>(P);
    DFK_CUDA(cudaGetLastError());
    ctx->launches += 2;
    unsigned long long nbad = 0;
    DFK_CUDA(cudaMemcpyAsync(&nbad, nbad_dev, sizeof(nbad), cudaMemcpyDeviceToHost, st));
    DFK_CUDA(cudaStreamSynchronize(st));
    if (nbad_out) *nbad_out = static_cast<int64_t>(nbad);
    return DFK_OK;
}

int dfk_text_release(dfk_ctx* ctx) {
    DFK_ENTER(ctx);
    DFK_CUDA(cudaStreamSynchronize(ctx->stream()));
    for (int i : {3, 4}) {
        if (ctx->post[i].ptr) cudaFree(ctx->post[i].ptr);
        ctx->post[i].ptr = nullptr;
        ctx->post[i].bytes = 0;
    }
    ctx->text_bytes = ctx->text_rows = ctx->text_chunks = 0;
    return DFK_OK;
}

int dfk_widen_dev(dfk_ctx* ctx, const void* src_dev, int32_t dtype, int64_t T, int64_t C, int32_t time_major, double scale,
                  double offset, double* out_dev, int64_t ld_c) {
    DFK_ENTER(ctx);
    if (!dtype_bytes(dtype)) return fail(DFK_ERR_ARG, "unknown sample type %d", dtype);
    if (T < 0 || C < 0 || ld_c < T) return fail(DFK_ERR_ARG, "bad geometry: T=%lld C=%lld ld_c=%lld", (long long)T, (long long)C, (long long)ld_c);
    if (T == 0 || C == 0) return DFK_OK;
    if (!src_dev || !out_dev) return fail(DFK_ERR_ARG, "null pointer");
    cudaStream_t st = ctx->stream();
    if (time_major && C > 1 && C <= dfk::kWidenFewMax) {
        const int grid = static_cast<int>(std::min<int64_t>((T + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 16));
        dfk::widen_few_channels_kernel<<<grid, 256, 0, st>>>(src_dev, dtype, T, static_cast<int>(C), scale, offset, out_dev, ld_c);
    } else if (time_major && C > 1) {
        const int64_t tiles = ((T + dfk::kWidenTile - 1) / dfk::kWidenTile) * ((C + dfk::kWidenTile - 1) / dfk::kWidenTile);
        const int grid = static_cast<int>(std::min<int64_t>(tiles, static_cast<int64_t>(ctx->sm_count) * 16));
        dfk::widen_time_major_kernel<<<grid, dfk::kWidenTile * 8, 0, st>>>(src_dev, dtype, T, C, scale, offset, out_dev, ld_c);
    } else {
        const int grid = static_cast<int>(std::min<int64_t>((T * C + 255) / 256, static_cast<int64_t>(ctx->sm_count) * 16));
        dfk::widen_channel_major_kernel<<<grid, 256, 0, st>>>(src_dev, dtype, T, C, T, scale, offset, out_dev, ld_c);
    }
    ctx->launches++;
    DFK_CUDA(cudaGetLastError());
    return DFK_OK;
}

static int ingest_binary(dfk_ctx* ctx, const ByteSource& src, int64_t base_off, int32_t dtype, int64_t T, int64_t C,
                         int32_t time_major, double scale, double offset, double* out_dev, int64_t ld_c) {
    const size_t es = dtype_bytes(dtype);
    if (!es) return fail(DFK_ERR_ARG, "unknown sample type %d", dtype);
    if (T < 0 || C < 0 || ld_c < T) return fail(DFK_ERR_ARG, "bad geometry: T=%lld C=%lld ld_c=%lld", (long long)T, (long long)C, (long long)ld_c);
    if (T == 0 || C == 0) return DFK_OK;
    if (!out_dev) return fail(DFK_ERR_ARG, "null output");
    const size_t slab_target = ctx->host_slab_bytes ? ctx->host_slab_bytes : (static_cast<size_t>(128) << 20);
    // a slab is a run of whole time steps: of all channels (time-major) or of one channel (channel-major)
    const int64_t step_bytes = static_cast<int64_t>(es) * (time_major ? C : 1);
    const int64_t steps_per_slab = std::max<int64_t>(1, std::min<int64_t>(T, static_cast<int64_t>(slab_target) / step_bytes));
    int rc = DFK_OK;
    for (int i = 0; i < 2 && !rc; ++i) rc = ensure(ctx, ctx->slab[i], static_cast<size_t>(steps_per_slab * step_bytes));
    if (rc) return rc;
    HostCallGuard hg(ctx);
    cudaStream_t st = ctx->stream();
    int64_t slab_index = 0;
    const int64_t outer = time_major ? 1 : C;
    for (int64_t c = 0; c < outer; ++c) {
        for (int64_t t0 = 0; t0 < T; t0 += steps_per_slab, ++slab_index) {
            const int s = static_cast<int>(slab_index & 1);
            const int64_t nt = std::min(steps_per_slab, T - t0);
            if (slab_index >= 2) DFK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->consumed[s], 0));
            const int64_t off = base_off + (time_major ? t0 * step_bytes : (c * T + t0) * static_cast<int64_t>(es));
            rc = source_to_device(ctx, src, off, ctx->slab[s].ptr, static_cast<size_t>(nt * step_bytes));
            if (rc) return rc;
            DFK_CUDA(cudaEventRecord(ctx->copied[s], ctx->copy_stream));
            DFK_CUDA(cudaStreamWaitEvent(st, ctx->copied[s], 0));
            if (time_major)
                rc = dfk_widen_dev(ctx, ctx->slab[s].ptr, dtype, nt, C, 1, scale, offset, out_dev + t0, ld_c);
            else
                rc = dfk_widen_dev(ctx, ctx->slab[s].ptr, dtype, nt, 1, 0, scale, offset, out_dev + c * ld_c + t0, nt);
            if (rc) return rc;
            DFK_CUDA(cudaEventRecord(ctx->consumed[s], st));
        }
    }
    DFK_CUDA(cudaStreamSynchronize(st));
    DFK_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    hg.done();
    return DFK_OK;
}

int dfk_ingest_binary_host(dfk_ctx* ctx, const void* src_host, int32_t dtype, int64_t T, int64_t C, int32_t time_major,
                           double scale, double offset, double* out_dev, int64_t ld_c) {
    DFK_ENTER(ctx);
    if (!src_host && T * C > 0) return fail(DFK_ERR_ARG, "null source");
    ByteSource src;
    src.host = static_cast<const char*>(src_host);
    src.pageable = src_host ? is_pageable(src_host) : true;
    return ingest_binary(ctx, src, 0, dtype, T, C, time_major, scale, offset, out_dev, ld_c);
}

int dfk_ingest_binary_file(dfk_ctx* ctx, const char* path, int64_t byte_offset, int32_t dtype, int64_t T, int64_t C,
                           int32_t time_major, double scale, double offset, double* out_dev, int64_t ld_c) {
    DFK_ENTER(ctx);
    if (!path) return fail(DFK_ERR_ARG, "null path");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(DFK_ERR_ARG, "cannot open %s: %s", path, std::strerror(errno));
    FileCloser closer{fd};
    struct stat sb;
    if (fstat(fd, &sb) != 0) return fail(DFK_ERR_ARG, "cannot stat %s", path);
    const int64_t need = byte_offset + T * C * static_cast<int64_t>(dtype_bytes(dtype));
    if (dtype_bytes(dtype) && static_cast<int64_t>(sb.st_size) < need)
        return fail(DFK_ERR_ARG, "%s holds %lld bytes, %lld needed", path, (long long)sb.st_size, (long long)need);
    ByteSource src;
    src.fd = fd;
    return ingest_binary(ctx, src, byte_offset, dtype, T, C, time_major, scale, offset, out_dev, ld_c);
}

}  // extern "C"
