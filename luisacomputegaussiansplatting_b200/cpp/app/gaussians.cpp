// gaussians.cpp -- binary/ascii PLY reader for INRIA 3DGS checkpoints + activations.
// Behaviour follows the reference's loader (app/gaussians.cpp:15-35 activations, :75-171 layout):
//   f_dc_c          -> feature[g][0][c]
//   f_rest_i        -> feature[g][1 + i % 15][i / 15]      (file is channel-major, memory is RGB-interleaved)
//   opacity         -> sigmoid,  scale_k -> exp,  rot_k -> normalised quaternion (r,x,y,z)
#include "gaussians.h"

#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <thread>
#include <unordered_set>

namespace lcgs
{

namespace
{
std::mutex                g_pinned_mutex;
std::unordered_set<void*> g_pinned;  // blocks that came from cudaHostAlloc (the rest came from malloc)
}  // namespace

void* host_alloc_pinned(size_t bytes)
{
    if (bytes == 0) bytes = 1;
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess && p) {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        g_pinned.insert(p);
        return p;
    }
    (void)cudaGetLastError();  // no device / out of pinned memory: pageable memory still works, just slower
    p = std::malloc(bytes);
    if (!p) throw std::bad_alloc();
    return p;
}

void host_free_pinned(void* p)
{
    if (!p) return;
    bool pinned;
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        pinned = g_pinned.erase(p) != 0;
    }
    if (pinned) cudaFreeHost(p);
    else std::free(p);
}

float GaussiansData::opacity_activation(float x) { return 1.0f / (1.0f + std::exp(-x)); }
float GaussiansData::scaling_activation(float x) { return std::exp(x); }
void  GaussiansData::rotation_activation(float& r, float& x, float& y, float& z)
{
    const float norm = std::sqrt(x * x + y * y + z * z + r * r);
    r /= norm; x /= norm; y /= norm; z /= norm;
}

void GaussiansData::resize(int N)
{
    num_gaussians  = N;
    const int feat = (sh_deg + 1) * (sh_deg + 1);
    pos.assign((size_t)N * 3, 0.f);
    feature.assign((size_t)N * feat * 3, 0.f);
    opacity.assign((size_t)N, 0.f);
    scale.assign((size_t)N * 3, 0.f);
    rotq.assign((size_t)N * 4, 0.f);
}

namespace
{
struct Prop {
    std::string name;
    int         size;    // bytes in a binary file
    bool        is_f32;
    int         offset;  // byte offset inside a vertex record
};

int type_size(const std::string& t, bool* is_f32)
{
    *is_f32 = (t == "float" || t == "float32");
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "int32" || t == "uint32" || *is_f32) return 4;
    if (t == "double" || t == "float64") return 8;
    return -1;
}

bool fail(std::string* err, const std::string& msg)
{
    if (err) *err = msg;
    return false;
}
}  // namespace

bool read_gs_ply(GaussiansData& gs, const std::filesystem::path& fpath, std::string* err, int threads)
{
    std::ifstream in(fpath, std::ios::binary);
    if (!in) return fail(err, "cannot open " + fpath.string());
    std::string line;
    if (!std::getline(in, line) || line.rfind("ply", 0) != 0) return fail(err, "not a PLY file");
    bool              binary = false, in_vertex = false, have_vertex = false;
    long              count = 0;
    std::vector<Prop> props;
    int               stride = 0;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string        tok;
        ls >> tok;
        if (tok == "format") {
            std::string f;
            ls >> f;
            if (f == "binary_little_endian") binary = true;
            else if (f == "ascii") binary = false;
            else return fail(err, "unsupported PLY format " + f);
        } else if (tok == "element") {
            std::string name;
            long        n;
            ls >> name >> n;
            in_vertex = (name == "vertex");
            if (in_vertex) {
                if (have_vertex) return fail(err, "duplicate vertex element");
                have_vertex = true;
                count       = n;
            } else if (!have_vertex) {
                return fail(err, "elements before 'vertex' are not supported");
            }
        } else if (tok == "property" && in_vertex) {
            std::string type, name;
            ls >> type;
            if (type == "list") return fail(err, "list property in vertex element");
            ls >> name;
            bool      f32;
            const int sz = type_size(type, &f32);
            if (sz < 0) return fail(err, "unknown property type " + type);
            props.push_back({ name, sz, f32, stride });
            stride += sz;
        } else if (tok == "end_header") {
            break;
        }
    }
    if (!have_vertex) return fail(err, "No vertex element in the ply file");
    if (count < 0 || count > 0x7FFFFFFF) return fail(err, "bad vertex count");

    const int feat = (gs.sh_deg + 1) * (gs.sh_deg + 1);
    // destination of each property: (array id, element stride, element offset)
    enum { POS, FEAT, OPA, SCA, ROT, NONE };
    struct Dst { int arr = NONE; int stride = 0; int off = 0; };
    std::vector<Dst>         dst(props.size());
    std::map<std::string, int> seen;
    for (size_t k = 0; k < props.size(); k++) {
        const std::string& n = props[k].name;
        Dst                d;
        if (n == "x" || n == "y" || n == "z") d = { POS, 3, n[0] - 'x' };
        else if (n.rfind("f_dc_", 0) == 0) {
            const int c = std::stoi(n.substr(5));
            if (c >= 0 && c < 3) d = { FEAT, feat * 3, c };
        } else if (n.rfind("f_rest_", 0) == 0) {
            const int i = std::stoi(n.substr(7));
            if (feat > 1 && i >= 0 && i < (feat - 1) * 3) d = { FEAT, feat * 3, (1 + i % (feat - 1)) * 3 + i / (feat - 1) };
        } else if (n == "opacity") d = { OPA, 1, 0 };
        else if (n.rfind("scale_", 0) == 0) {
            const int c = std::stoi(n.substr(6));
            if (c >= 0 && c < 3) d = { SCA, 3, c };
        } else if (n.rfind("rot_", 0) == 0) {
            const int c = std::stoi(n.substr(4));
            if (c >= 0 && c < 4) d = { ROT, 4, c };
        }
        if (d.arr != NONE) {
            if (!props[k].is_f32) return fail(err, "property " + n + " is not float32");
            seen[n] = 1;
        }
        dst[k] = d;
    }
    const size_t need = 3 + 3 + (size_t)(feat - 1) * 3 + 1 + 3 + 4;
    if (seen.size() != need) return fail(err, "missing Gaussian properties (found " + std::to_string(seen.size()) + " of " +
                                                   std::to_string(need) + ")");
    gs.resize((int)count);
    float* arrays[5] = { gs.pos.data(), gs.feature.data(), gs.opacity.data(), gs.scale.data(), gs.rotq.data() };

    auto activate = [&](long g0, long g1) {
        for (long g = g0; g < g1; g++) {
            gs.opacity[g] = GaussiansData::opacity_activation(gs.opacity[g]);
            for (int c = 0; c < 3; c++) gs.scale[3 * g + c] = GaussiansData::scaling_activation(gs.scale[3 * g + c]);
            GaussiansData::rotation_activation(gs.rotq[4 * g], gs.rotq[4 * g + 1], gs.rotq[4 * g + 2], gs.rotq[4 * g + 3]);
        }
    };
    if (binary) {
        // map the body; every worker de-interleaves and activates its own range of rows
        const std::streamoff body = in.tellg();
        in.close();
        const size_t need_bytes = (size_t)count * (size_t)stride;
        const int    fd         = ::open(fpath.c_str(), O_RDONLY);
        if (fd < 0) return fail(err, "cannot open " + fpath.string());
        struct stat st;
        if (fstat(fd, &st) != 0 || (size_t)st.st_size < (size_t)body + need_bytes) {
            ::close(fd);
            return fail(err, "truncated PLY body");
        }
        const size_t map_len = (size_t)body + need_bytes;
        void*        map     = map_len ? mmap(nullptr, map_len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
        ::close(fd);
        if (map_len && map == MAP_FAILED) return fail(err, "mmap failed");
        if (map_len) madvise(map, map_len, MADV_SEQUENTIAL);
        const char* base = static_cast<const char*>(map) + body;
        int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
        nt     = std::max(1, std::min(nt, (int)((count + 65535) / 65536)));
        auto work = [&](long g0, long g1) {
            for (long g = g0; g < g1; g++) {
                const char* row = base + (size_t)g * stride;
                for (size_t k = 0; k < props.size(); k++)
                    if (dst[k].arr != NONE)
                        std::memcpy(arrays[dst[k].arr] + (size_t)g * dst[k].stride + dst[k].off, row + props[k].offset, 4);
            }
            activate(g0, g1);
        };
        std::vector<std::thread> pool;
        const long               per = (count + nt - 1) / nt;
        for (int t = 1; t < nt; t++) pool.emplace_back(work, std::min<long>(count, t * per), std::min<long>(count, (t + 1) * per));
        work(0, std::min<long>(count, per));
        for (auto& th : pool) th.join();
        if (map_len) munmap(map, map_len);
    } else {
        for (long g = 0; g < count; g++)
            for (size_t k = 0; k < props.size(); k++) {
                double v;
                if (!(in >> v)) return fail(err, "truncated ascii PLY body");
                if (dst[k].arr != NONE) arrays[dst[k].arr][(size_t)g * dst[k].stride + dst[k].off] = (float)v;
            }
        activate(0, count);
    }
    return true;
}

bool write_gs_ply(const std::filesystem::path& fpath, int P, const float* pos, const float* sh, const float* logit_opacity,
                  const float* log_scale, const float* raw_rot)
{
    std::ofstream out(fpath, std::ios::binary);
    if (!out) return false;
    out << "ply\nformat binary_little_endian 1.0\nelement vertex " << P << "\n";
    const char* xyz[3] = { "x", "y", "z" };
    for (auto n : xyz) out << "property float " << n << "\n";
    for (auto n : { "nx", "ny", "nz" }) out << "property float " << n << "\n";
    for (int c = 0; c < 3; c++) out << "property float f_dc_" << c << "\n";
    for (int i = 0; i < 45; i++) out << "property float f_rest_" << i << "\n";
    out << "property float opacity\n";
    for (int c = 0; c < 3; c++) out << "property float scale_" << c << "\n";
    for (int c = 0; c < 4; c++) out << "property float rot_" << c << "\n";
    out << "end_header\n";
    std::vector<float> row(3 + 3 + 3 + 45 + 1 + 3 + 4);
    for (int g = 0; g < P; g++) {
        size_t k = 0;
        for (int c = 0; c < 3; c++) row[k++] = pos[3 * g + c];
        for (int c = 0; c < 3; c++) row[k++] = 0.f;
        for (int c = 0; c < 3; c++) row[k++] = sh[(size_t)g * 48 + c];
        for (int i = 0; i < 45; i++) row[k++] = sh[(size_t)g * 48 + (1 + i % 15) * 3 + i / 15];
        row[k++] = logit_opacity[g];
        for (int c = 0; c < 3; c++) row[k++] = log_scale[3 * g + c];
        for (int c = 0; c < 4; c++) row[k++] = raw_rot[4 * g + c];
        out.write(reinterpret_cast<const char*>(row.data()), (std::streamsize)(row.size() * sizeof(float)));
    }
    return (bool)out;
}

}  // namespace lcgs
