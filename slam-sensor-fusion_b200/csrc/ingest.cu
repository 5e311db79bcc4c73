// ingest.cu -- the data formats on either side of the registration path, natively, and
// the map kept resident in HBM so that a re-crop is a window change, not a re-upload (row N2).
//
//   ssf_pcd_read              binary / ascii PCD (the recorder's tiles, reference
//                             mapping/src/map_data_save_node.cpp:71-80,101-113; read back at
//                             localization/src/global_map_frames_manager.cpp:101,129) -> packed xyz
//   ssf_cloud_from_pointcloud2  pcl::fromROSMsg (localization_node.cpp:290-291,
//                             map_data_save_node.cpp:66) for the xyz float32 fields of a
//                             sensor_msgs/PointCloud2 byte buffer -> float4 on the device
//   ssf_map_*                 GlobalMapFramesManager::getMapCloud / mergeScansAndSave
//                             (global_map_frames_manager.cpp:93-151): tiles -> pinned host -> HBM,
//                             concatenated, pcl::VoxelGrid on the device; the result STAYS in HBM, and
//                             cropPointCloudThroughRadius + setTargetPointCloud
//                             (localization_node.cpp:300-305) become a device-side crop + index build.
//
// File parsing is host code; every byte of geometry work runs on the device.
#include <dirent.h>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "ingest.cuh"
#include "map_index.cuh"
#include "preprocess.cuh"
#include "voxel_grid.cuh"

namespace ssf {

// ---- PCD -------------------------------------------------------------------------------------------
struct PcdHeader {
    size_t n_points = 0, point_bytes = 0, data_offset = 0;
    size_t off[3] = {0, 0, 0};      // byte offset of x, y, z inside a record (binary)
    int size[3] = {4, 4, 4};        // SIZE of x, y, z
    char type[3] = {'F', 'F', 'F'}; // TYPE of x, y, z
    int col[3] = {0, 1, 2};         // column of x, y, z (ascii)
    int n_cols = 0;
    bool binary = true;
};

static std::vector<std::string> split_ws(const std::string &s)
{
    std::vector<std::string> out;
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && isspace((unsigned char)s[i])) ++i;
        size_t j = i;
        while (j < s.size() && !isspace((unsigned char)s[j])) ++j;
        if (j > i) out.push_back(s.substr(i, j - i));
        i = j;
    }
    return out;
}

static int parse_pcd_header(FILE *f, const char *path, PcdHeader &h)
{
    std::vector<std::string> fields, sizes, types, counts;
    long long width = -1, height = 1, points = -1;
    char line[4096];
    bool have_data = false;
    while (fgets(line, sizeof(line), f)) {
        std::string t(line);
        while (!t.empty() && (t.back() == '\n' || t.back() == '\r')) t.pop_back();
        if (t.empty() || t[0] == '#') continue;
        std::vector<std::string> w = split_ws(t);
        if (w.empty()) continue;
        std::string key = w[0];
        std::transform(key.begin(), key.end(), key.begin(), ::toupper);
        w.erase(w.begin());
        if (key == "FIELDS" || key == "COLUMNS") fields = w;
        else if (key == "SIZE") sizes = w;
        else if (key == "TYPE") types = w;
        else if (key == "COUNT") counts = w;
        else if (key == "WIDTH" && !w.empty()) width = atoll(w[0].c_str());
        else if (key == "HEIGHT" && !w.empty()) height = atoll(w[0].c_str());
        else if (key == "POINTS" && !w.empty()) points = atoll(w[0].c_str());
        else if (key == "DATA") {
            std::string kind = w.empty() ? "" : w[0];
            std::transform(kind.begin(), kind.end(), kind.begin(), ::tolower);
            if (kind == "binary") h.binary = true;
            else if (kind == "ascii") h.binary = false;
            else {
                set_error("%s: DATA %s is not supported (the recorder writes binary)", path, kind.c_str());
                return SSF_ERR_INVALID;
            }
            have_data = true;
            break;
        }
    }
    if (!have_data || fields.empty() || sizes.size() != fields.size() || types.size() != fields.size()) {
        set_error("%s: not a PCD file (FIELDS / SIZE / TYPE / DATA missing or inconsistent)", path);
        return SSF_ERR_INVALID;
    }
    if (counts.empty()) counts.assign(fields.size(), "1");
    if (counts.size() != fields.size()) {
        set_error("%s: COUNT does not match FIELDS", path);
        return SSF_ERR_INVALID;
    }
    if (points < 0) points = width >= 0 ? width * height : -1;
    if (points < 0) {
        set_error("%s: neither POINTS nor WIDTH given", path);
        return SSF_ERR_INVALID;
    }
    h.n_points = (size_t)points;
    size_t off = 0;
    int col = 0, found = 0;
    for (size_t i = 0; i < fields.size(); ++i) {
        const int sz = atoi(sizes[i].c_str()), cnt = std::max(1, atoi(counts[i].c_str()));
        for (int k = 0; k < 3; ++k)
            if (fields[i] == (k == 0 ? "x" : k == 1 ? "y" : "z")) {
                h.off[k] = off;
                h.size[k] = sz;
                h.type[k] = types[i].empty() ? 'F' : (char)toupper(types[i][0]);
                h.col[k] = col;
                ++found;
            }
        off += (size_t)sz * cnt;
        col += cnt;
    }
    if (found != 3) {
        set_error("%s: fields x, y, z not all present", path);
        return SSF_ERR_INVALID;
    }
    for (int k = 0; k < 3; ++k)
        if (h.binary && !((h.type[k] == 'F' && (h.size[k] == 4 || h.size[k] == 8)))) {
            set_error("%s: x / y / z must be float32 or float64", path);
            return SSF_ERR_INVALID;
        }
    h.point_bytes = off;
    h.n_cols = col;
    h.data_offset = (size_t)ftell(f);
    return SSF_OK;
}

// records of a binary PCD (or a PointCloud2 buffer) -> float4 (w = 1); x / y / z float32 or float64 at
// given byte offsets.  One thread per point; the 12..32-byte records are read through the read-only
// path, the 16-byte stores are coalesced.
__global__ void __launch_bounds__(256)
    extract_xyz_kernel(const unsigned char *__restrict__ raw, size_t n, size_t point_bytes, uint32_t ox, uint32_t oy,
                       uint32_t oz, int f64, int swap, float4 *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char *p = raw + i * point_bytes;
    float v[3];
    const uint32_t o[3] = {ox, oy, oz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (f64) {
            unsigned long long b = 0;
            for (int j = 0; j < 8; ++j) b |= (unsigned long long)p[o[k] + (swap ? 7 - j : j)] << (8 * j);
            v[k] = (float)__longlong_as_double((long long)b);
        } else {
            uint32_t b = 0;
            for (int j = 0; j < 4; ++j) b |= (uint32_t)p[o[k] + (swap ? 3 - j : j)] << (8 * j);
            v[k] = __uint_as_float(b);
        }
    }
    out[i] = make_float4(v[0], v[1], v[2], 1.0f);
}

int extract_xyz_device(const unsigned char *raw_dev, size_t n, size_t point_bytes, size_t ox, size_t oy, size_t oz,
                       bool f64, bool swap, float4 *out, cudaStream_t st)
{
    if (n == 0) return SSF_OK;
    extract_xyz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(raw_dev, n, point_bytes, (uint32_t)ox, (uint32_t)oy,
                                                                   (uint32_t)oz, f64 ? 1 : 0, swap ? 1 : 0, out);
    SSF_LAUNCHED();
    return SSF_OK;
}

// bytes the records of a file take in a host buffer (binary: the file's own records; ascii: float32 x y z)
static size_t pcd_record_bytes(const PcdHeader &h) { return h.n_points * (h.binary ? h.point_bytes : 12); }

static int open_pcd(const char *path, PcdHeader &h, FILE **f_out)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_error("%s: %s", path, strerror(errno));
        return SSF_ERR_INVALID;
    }
    const int rc = parse_pcd_header(f, path, h);
    if (rc != SSF_OK) {
        fclose(f);
        return rc;
    }
    *f_out = f;
    return SSF_OK;
}

// the data section of an opened file -> dst (pcd_record_bytes(h) bytes; may be pinned memory); closes f.
// After an ascii file h describes packed float32 x y z records.
static int read_pcd_body(const char *path, PcdHeader &h, FILE *f, unsigned char *dst)
{
    if (h.binary) {
        const size_t want = h.n_points * h.point_bytes;
        const size_t got = want ? fread(dst, 1, want, f) : 0;
        fclose(f);
        if (got != want) {
            set_error("%s: truncated (%zu of %zu data bytes)", path, got, want);
            return SSF_ERR_INVALID;
        }
        return SSF_OK;
    }
    float *out = reinterpret_cast<float *>(dst);  // ascii: one point per line, n_cols numbers
    std::vector<double> row((size_t)h.n_cols);
    for (size_t i = 0; i < h.n_points; ++i) {
        for (int c = 0; c < h.n_cols; ++c)
            if (fscanf(f, "%lf", &row[(size_t)c]) != 1) {
                fclose(f);
                set_error("%s: ascii data ends at point %zu", path, i);
                return SSF_ERR_INVALID;
            }
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (float)row[(size_t)h.col[k]];
    }
    fclose(f);
    h.point_bytes = 12;
    h.off[0] = 0; h.off[1] = 4; h.off[2] = 8;
    h.size[0] = h.size[1] = h.size[2] = 4;
    return SSF_OK;
}

static int read_pcd_records(const char *path, PcdHeader &h, std::vector<unsigned char> &rec)
{
    FILE *f = nullptr;
    SSF_TRY(open_pcd(path, h, &f));
    rec.resize(pcd_record_bytes(h));
    return read_pcd_body(path, h, f, rec.data());
}

}  // namespace ssf

using namespace ssf;

// ---- handles ---------------------------------------------------------------------------------------
struct ssf_map {
    ssf_ctx_ref ctx;
    PreprocWork w;       // w.in = the resident cloud (float4), w.out / w.idx = the last crop
    size_t n = 0;
    uint32_t last_crop = 0;
    double ingest_ms = 0.0;  // stream time of the last tiles -> HBM -> merge (includes waiting for the host's file reads)
    double merge_ms = 0.0;   // ... of which the voxel filter over the concatenated cloud
};

#define ING_ARG(cond, msg)                 \
    do {                                   \
        if (!(cond)) {                     \
            ssf::set_error("%s", msg);     \
            return SSF_ERR_INVALID;        \
        }                                  \
    } while (0)

extern "C" int ssf_pcd_read(const char *path, float *xyz_out, size_t cap_points, size_t *n_points)
{
    ING_ARG(path && n_points, "ssf_pcd_read: NULL argument");
    PcdHeader h;
    std::vector<unsigned char> rec;
    SSF_TRY(read_pcd_records(path, h, rec));
    *n_points = h.n_points;
    if (!xyz_out) return SSF_OK;  // size query
    ING_ARG(cap_points >= h.n_points, "ssf_pcd_read: output buffer too small");
    for (size_t i = 0; i < h.n_points; ++i) {
        const unsigned char *p = rec.data() + i * h.point_bytes;
        for (int k = 0; k < 3; ++k) {
            if (h.size[k] == 8) {
                double d;
                memcpy(&d, p + h.off[k], 8);
                xyz_out[3 * i + k] = (float)d;
            } else {
                memcpy(&xyz_out[3 * i + k], p + h.off[k], 4);
            }
        }
    }
    return SSF_OK;
}

extern "C" int ssf_pcd_write_binary(const char *path, const float *xyz, size_t n, size_t stride_bytes)
{
    ING_ARG(path && (xyz || n == 0), "ssf_pcd_write_binary: NULL argument");
    ING_ARG(stride_bytes >= 12 && stride_bytes % 4 == 0, "stride_bytes must be a multiple of 4 and >= 12");
    FILE *f = fopen(path, "wb");
    if (!f) {
        set_error("%s: %s", path, strerror(errno));
        return SSF_ERR_INVALID;
    }
    // the layout pcl::io::savePCDFileBinary gives a PointXYZ cloud: FIELDS x y z, 12 bytes per point
    fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\n"
               "COUNT 1 1 1\nWIDTH %zu\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %zu\nDATA binary\n", n, n);
    for (size_t i = 0; i < n; ++i)
        fwrite(reinterpret_cast<const unsigned char *>(xyz) + i * stride_bytes, 4, 3, f);
    const bool ok = fclose(f) == 0;
    if (!ok) {
        set_error("%s: write failed", path);
        return SSF_ERR_INVALID;
    }
    return SSF_OK;
}

extern "C" int ssf_cloud_from_pointcloud2(ssf_ctx *ctx, const unsigned char *data, size_t n_points, size_t point_step,
                                          size_t off_x, size_t off_y, size_t off_z, int is_bigendian, float *xyz_out)
{
    ING_ARG(ctx && (n_points == 0 || (data && xyz_out)), "ssf_cloud_from_pointcloud2: NULL argument");
    ING_ARG(point_step >= 12 && off_x + 4 <= point_step && off_y + 4 <= point_step && off_z + 4 <= point_step,
            "ssf_cloud_from_pointcloud2: field offsets outside the point step");
    if (n_points == 0) return SSF_OK;
    ssf_ctx_ref c(ctx);
    SSF_TRY(c.use());
    DevBuf<unsigned char> raw;
    DevBuf<float4> out;
    SSF_TRY(raw.reserve(n_points * point_step));
    SSF_TRY(out.reserve(n_points));
    SSF_CUDA(cudaMemcpyAsync(raw.p, data, n_points * point_step, cudaMemcpyHostToDevice, c.stream()));
    SSF_TRY(extract_xyz_device(raw.p, n_points, point_step, off_x, off_y, off_z, false, is_bigendian != 0, out.p, c.stream()));
    SSF_CUDA(cudaMemcpyAsync(xyz_out, out.p, n_points * sizeof(float4), cudaMemcpyDeviceToHost, c.stream()));
    SSF_CUDA(cudaStreamSynchronize(c.stream()));
    return SSF_OK;
}

// ---- resident map ----------------------------------------------------------------------------------
extern "C" int ssf_map_create(ssf_ctx *ctx, const float *xyz, size_t n, size_t stride_bytes, ssf_map **out)
{
    ING_ARG(ctx && out && (xyz || n == 0), "ssf_map_create: NULL argument");
    ING_ARG(n < ((size_t)1 << 31), "ssf_map_create: more than 2^31 - 1 points");
    *out = nullptr;
    ssf_map *m = new (std::nothrow) ssf_map{ssf_ctx_ref(ctx)};
    if (!m) return SSF_ERR_NOMEM;
    int rc = m->ctx.use();
    if (rc == SSF_OK) rc = m->w.in.reserve(n ? n : 1);
    if (rc == SSF_OK) rc = m->w.out.reserve(n ? n : 1);
    if (rc == SSF_OK) rc = m->ctx.upload_cloud(xyz, n, stride_bytes, m->w.in.p);
    if (rc == SSF_OK && cudaStreamSynchronize(m->ctx.stream()) != cudaSuccess) rc = SSF_ERR_CUDA;
    if (rc != SSF_OK) {
        delete m;
        return rc;
    }
    m->n = n;
    *out = m;
    return SSF_OK;
}

extern "C" void ssf_map_destroy(ssf_map *m)
{
    if (!m) return;
    m->ctx.use();
    cudaStreamSynchronize(m->ctx.stream());
    delete m;
}

extern "C" size_t ssf_map_size(const ssf_map *m) { return m ? m->n : 0; }
extern "C" double ssf_map_ingest_ms(const ssf_map *m) { return m ? m->ingest_ms : 0.0; }
extern "C" double ssf_map_merge_ms(const ssf_map *m) { return m ? m->merge_ms : 0.0; }

// getMapCloud / mergeScansAndSave (global_map_frames_manager.cpp:93-151)
extern "C" int ssf_map_from_pcd_folder(ssf_ctx *ctx, const char *data_folder, const char *map_name, float voxel_size,
                                       int save, ssf_map **out)
{
    ING_ARG(ctx && data_folder && map_name && out, "ssf_map_from_pcd_folder: NULL argument");
    *out = nullptr;
    ssf_ctx_ref c(ctx);
    SSF_TRY(c.use());
    const std::string folder(data_folder), cached = folder + "/" + map_name + ".pcd";
    std::vector<std::string> files;
    bool merged = false;
    if (FILE *t = fopen(cached.c_str(), "rb")) {  // :97-102: the cached map is loaded as it is (no voxel filter)
        fclose(t);
        files.push_back(cached);
    } else {
        DIR *dir = opendir(folder.c_str());
        if (!dir) {
            set_error("Could not open DATA directory %s: %s", folder.c_str(), strerror(errno));  // :137-140
            return SSF_ERR_INVALID;
        }
        while (dirent *ent = readdir(dir)) {  // readdir order, like the reference
            const std::string name = ent->d_name;
            if (name.size() > 4 && name.substr(name.size() - 4) == ".pcd") files.push_back(folder + "/" + name);
        }
        closedir(dir);
        merged = true;
    }
    // tiles -> pinned host -> HBM: every tile's records are read straight into pinned memory and extracted
    // to float4 at the tile's offset of ONE cloud
    std::vector<PcdHeader> heads(files.size());
    std::vector<FILE *> fps(files.size(), nullptr);
    size_t total = 0, max_bytes = 0;
    auto close_all = [&] { for (FILE *f : fps) if (f) fclose(f); };
    for (size_t i = 0; i < files.size(); ++i) {
        const int rc0 = open_pcd(files[i].c_str(), heads[i], &fps[i]);
        if (rc0 != SSF_OK) {
            close_all();
            return rc0;
        }
        total += heads[i].n_points;
        max_bytes = std::max(max_bytes, pcd_record_bytes(heads[i]));
    }
    if (!(total < ((size_t)1 << 31))) close_all();
    ING_ARG(total < ((size_t)1 << 31), "ssf_map_from_pcd_folder: more than 2^31 - 1 points");
    ssf_map *m = new (std::nothrow) ssf_map{ssf_ctx_ref(ctx)};
    if (!m) {
        close_all();
        return SSF_ERR_NOMEM;
    }
    auto bail = [&](int rc) { close_all(); delete m; return rc; };
    VoxelWork vw;
    DevBuf<float4> &cat = merged ? vw.in : m->w.in;
    int rc = cat.reserve(total ? total : 1);
    if (rc != SSF_OK) return bail(rc);
    PinnedBuf<unsigned char> pin[2];
    DevBuf<unsigned char> stage[2];
    cudaEvent_t done[2] = {nullptr, nullptr}, e0 = nullptr, e1 = nullptr, ev = nullptr;
    for (int k = 0; k < 2; ++k) {
        if ((rc = pin[k].reserve(max_bytes ? max_bytes : 1)) != SSF_OK) return bail(rc);
        if ((rc = stage[k].reserve(max_bytes ? max_bytes : 1)) != SSF_OK) return bail(rc);
        cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
    }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&ev);
    cudaStream_t st = c.stream();
    cudaEventRecord(e0, st);
    size_t at = 0;
    for (size_t i = 0; i < files.size() && rc == SSF_OK; ++i) {  // double-buffered: the copy of tile i overlaps the staging of tile i + 1
        const int k = (int)(i & 1);
        if (i >= 2) cudaEventSynchronize(done[k]);
        PcdHeader &h = heads[i];
        rc = read_pcd_body(files[i].c_str(), h, fps[i], pin[k].p);
        fps[i] = nullptr;
        if (rc != SSF_OK) break;
        if (h.n_points) {
            if (cudaMemcpyAsync(stage[k].p, pin[k].p, h.n_points * h.point_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = SSF_ERR_CUDA;
            if (rc == SSF_OK)
                rc = extract_xyz_device(stage[k].p, h.n_points, h.point_bytes, h.off[0], h.off[1], h.off[2], h.size[0] == 8, false,
                                        cat.p + at, st);
        }
        cudaEventRecord(done[k], st);
        at += h.n_points;
    }
    uint32_t n_out = (uint32_t)total;
    cudaEventRecord(ev, st);
    if (rc == SSF_OK && merged && total > 0) {  // :143-146 pcl::VoxelGrid(voxel_size)
        int refused = 0;
        rc = vw.out.reserve(total);
        if (rc == SSF_OK) rc = voxel_downsample_device(vw, total, voxel_size, c.scratch(), st, &n_out, &refused);
        if (rc == SSF_OK) rc = m->w.in.reserve(n_out ? n_out : 1);
        if (rc == SSF_OK && n_out &&
            cudaMemcpyAsync(m->w.in.p, vw.out.p, (size_t)n_out * sizeof(float4), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            rc = SSF_ERR_CUDA;
    }
    cudaEventRecord(e1, st);
    if (rc == SSF_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = SSF_ERR_CUDA;
    float ms = 0.f, ms_merge = 0.f;
    if (rc == SSF_OK) cudaEventElapsedTime(&ms, e0, e1);
    if (rc == SSF_OK) cudaEventElapsedTime(&ms_merge, ev, e1);
    for (int k = 0; k < 2; ++k) cudaEventDestroy(done[k]);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaEventDestroy(ev);
    if (rc != SSF_OK) {
        if (rc == SSF_ERR_CUDA) set_error("ssf_map_from_pcd_folder: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
        return bail(rc);
    }
    m->n = n_out;
    m->ingest_ms = ms;
    m->merge_ms = ms_merge;
    if ((rc = m->w.out.reserve(m->n ? m->n : 1)) != SSF_OK) return bail(rc);
    if (merged && save) {  // :148 savePCDFileBinary(<folder>/<map_name>.pcd)
        std::vector<float> host((size_t)n_out * 4);
        if (n_out && cudaMemcpy(host.data(), m->w.in.p, (size_t)n_out * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess)
            return bail(SSF_ERR_CUDA);
        if ((rc = ssf_pcd_write_binary(cached.c_str(), host.data(), n_out, 16)) != SSF_OK) return bail(rc);
    }
    *out = m;
    return SSF_OK;
}

extern "C" int ssf_map_download(ssf_map *m, float *xyz_out, size_t cap_points)
{
    ING_ARG(m && (xyz_out || m->n == 0), "ssf_map_download: NULL argument");
    ING_ARG(cap_points >= m->n, "ssf_map_download: output buffer too small");
    SSF_TRY(m->ctx.use());
    if (m->n) SSF_CUDA(cudaMemcpy(xyz_out, m->w.in.p, m->n * sizeof(float4), cudaMemcpyDeviceToHost));
    return SSF_OK;
}

// applyUniformSubsample(map_cloud_, step) at node start-up (localization_node.cpp:20), in place in HBM
extern "C" int ssf_map_subsample(ssf_map *m, size_t point_step)
{
    ING_ARG(m, "ssf_map_subsample: m == NULL");
    SSF_TRY(m->ctx.use());
    uint32_t cnt = 0;
    SSF_TRY(subsample_device(m->w, m->n, point_step, &cnt, m->ctx.stream()));
    if (cnt) SSF_CUDA(cudaMemcpyAsync(m->w.in.p, m->w.out.p, (size_t)cnt * sizeof(float4), cudaMemcpyDeviceToDevice, m->ctx.stream()));
    SSF_CUDA(cudaStreamSynchronize(m->ctx.stream()));
    m->n = cnt;
    return SSF_OK;
}

static int crop_resident(ssf_map *m, const float center[3], double radius)
{
    SSF_TRY(m->ctx.use());
    uint32_t cnt = 0;
    SSF_TRY(crop_radius_device(m->w, m->n, center, radius, m->ctx.scratch(), &cnt, m->ctx.stream()));
    m->last_crop = cnt;
    return SSF_OK;
}

// cropPointCloudThroughRadius(map_T_sensor_, radius, map_cloud_, cropped) (localization_node.cpp:302) on the
// resident map; the cropped cloud goes to the host (debug topics, :359-372) and/or becomes the target
extern "C" int ssf_map_crop_radius(ssf_map *m, const float center[3], double radius, float *xyz_out, size_t cap_points,
                                   size_t *n_out, int32_t *indices_out)
{
    ING_ARG(m && center && n_out, "ssf_map_crop_radius: NULL argument");
    ING_ARG(radius >= 0.0, "ssf_map_crop_radius: radius < 0");
    SSF_TRY(crop_resident(m, center, radius));
    *n_out = m->last_crop;
    if (!xyz_out && !indices_out) return SSF_OK;
    ING_ARG(cap_points >= m->last_crop, "ssf_map_crop_radius: output buffer too small");
    cudaStream_t st = m->ctx.stream();
    if (m->last_crop && xyz_out)
        SSF_CUDA(cudaMemcpyAsync(xyz_out, m->w.out.p, (size_t)m->last_crop * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (m->last_crop && indices_out)
        SSF_CUDA(cudaMemcpyAsync(indices_out, m->w.idx.p, (size_t)m->last_crop * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SSF_CUDA(cudaStreamSynchronize(st));
    return SSF_OK;
}

// crop + icp_->setTargetPointCloud(ref_cropped_map_cloud_) (localization_node.cpp:300-305) with no host copy
extern "C" int ssf_map_crop_to_target(ssf_map *m, ssf_icp *icp, const float center[3], double radius, size_t *n_out)
{
    ING_ARG(m && icp && center, "ssf_map_crop_to_target: NULL argument");
    ING_ARG(radius >= 0.0, "ssf_map_crop_to_target: radius < 0");
    SSF_TRY(crop_resident(m, center, radius));
    if (n_out) *n_out = m->last_crop;
    return icp_set_target_device(icp, m->w.out.p, m->last_crop, m->ctx);
}
