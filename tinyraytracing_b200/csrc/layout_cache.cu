// Layout cache: the acceleration layouts trt_scene_create builds from a scene description (reference-topology nodes,
// the 4-wide fast layout with its triangle order, tie keys, reference-leaf boxes, light boxes) written to / read from
// one file, so that a scene that was created once is created again without buildAccel / buildWide — on the
// 10 M-triangle stress mesh those are 17 of trt_scene_create's seconds, most of them the sequential insertion-based
// optimisation of the tree.  Host code only; nothing here touches the device.
//
// The file is keyed by a hash of everything the builders read — the description's arrays and counts, the environment
// switches that shape the layout, the record sizes and a format version — and carries a checksum of its payload.  A file
// that is missing, foreign, stale, truncated or corrupt is never an error: the layouts are built as usual and the file
// is rewritten (beside the target, then renamed over it).
#include "accel.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <unistd.h>

namespace trt
{
namespace
{
constexpr uint32_t kFormatVersion = 1;
constexpr char kMagic[8] = {'T', 'R', 'T', 'L', 'A', 'Y', 'O', 'T'};

uint64_t fnv(const void *data, size_t n, uint64_t h)
{
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < n; ++i)
        h = (h ^ p[i]) * 1099511628211ull;
    return h;
}
template <typename T>
uint64_t fnvValue(const T &v, uint64_t h)
{
    static_assert(std::is_trivially_copyable<T>::value, "POD");
    return fnv(&v, sizeof v, h);
}

struct Header
{
    char magic[8];
    uint32_t version, use_wide;
    uint64_t key, payload_bytes, checksum;
};

// the scalar part of AccelBuild, in one record (explicitly sized fields, no padding surprises: all 4 / 8 bytes, 8-aligned)
struct Scalars
{
    int32_t root_link, n_leaves, ref_depth, root_is_reference_leaf, n_sliver, n_needle, wide_root, wide_depth;
    uint32_t miss_rank;
    float scene_scale;
    double sah_ref, sah_wide;
};

struct Writer
{
    std::string buf;
    template <typename T>
    void vec(const std::vector<T> &v)
    {
        static_assert(std::is_trivially_copyable<T>::value, "POD");
        const uint64_t n = v.size();
        buf.append(reinterpret_cast<const char *>(&n), sizeof n);
        if (n)
            buf.append(reinterpret_cast<const char *>(v.data()), n * sizeof(T));
    }
};
struct Reader
{
    const char *p, *end;
    bool ok = true;
    template <typename T>
    void vec(std::vector<T> &v)
    {
        uint64_t n = 0;
        if (!ok || (size_t)(end - p) < sizeof n)
        {
            ok = false;
            return;
        }
        std::memcpy(&n, p, sizeof n);
        p += sizeof n;
        if (n > (uint64_t)(end - p) / sizeof(T))
        {
            ok = false;
            return;
        }
        v.resize((size_t)n);
        if (n)
            std::memcpy(v.data(), p, (size_t)n * sizeof(T));
        p += (size_t)n * sizeof(T);
    }
};

template <typename IO>
void allVectors(IO &io, AccelBuild &ab)
{
    io.vec(ab.ref_nodes), io.vec(ab.tri_geom), io.vec(ab.tri_key), io.vec(ab.tri_rank), io.vec(ab.rank_tri);
    io.vec(ab.wide_nodes), io.vec(ab.fast_geom), io.vec(ab.fast_key), io.vec(ab.fast_rank), io.vec(ab.fast_orig);
    io.vec(ab.fast_leaf), io.vec(ab.ref_leaf_box), io.vec(ab.light_box), io.vec(ab.ref_leaf_parent);
}
} // namespace

uint64_t layoutKey(const trt_scene_desc &d)
{
    uint64_t h = 1469598103934665603ull;
    h = fnvValue(kFormatVersion, h);
    const uint32_t library = TRT_VERSION; // a layout written by another version of the builders is rebuilt, not trusted
    h = fnvValue(library, h);
    const uint32_t sizes[4] = {(uint32_t)sizeof(RefNode), (uint32_t)sizeof(WideNode), (uint32_t)sizeof(TriGeom), (uint32_t)TRT_WIDE_STACK};
    h = fnv(sizes, sizeof sizes, h);
    // the switches that shape the layout
    for (const char *name : {"TRT_WIDE_SOURCE", "TRT_FAST_LEAF", "TRT_REINSERT", "TRT_COLLAPSE"})
    {
        const char *e = getenv(name);
        h = fnv(name, std::strlen(name) + 1, h);
        if (e)
            h = fnv(e, std::strlen(e) + 1, h);
    }
    // what buildAccel / buildWide read of the description
    const int32_t counts[4] = {d.n_tris, d.n_nodes, d.n_materials, d.n_lights};
    h = fnv(counts, sizeof counts, h);
    const size_t nt = (size_t)(d.n_tris > 0 ? d.n_tris : 0), nn = (size_t)(d.n_nodes > 0 ? d.n_nodes : 0);
    if (nt)
    {
        h = fnv(d.v, nt * 9 * sizeof(float), h);
        h = fnv(d.normal, nt * 3 * sizeof(float), h);
        h = fnv(d.mtl, nt * sizeof(int32_t), h);
    }
    if (nn)
    {
        h = fnv(d.node_box, nn * 6 * sizeof(float), h);
        h = fnv(d.node_link, nn * 4 * sizeof(int32_t), h);
    }
    for (int i = 0; i < d.n_materials; ++i)
        h = fnvValue(d.materials[i].is_emissive, h);
    for (int i = 0; i < d.n_lights; ++i)
        h = fnvValue(d.lights[i].material, h);
    return h;
}

std::string saveLayout(const AccelBuild &ab_in, bool use_wide, uint64_t key, const char *path)
{
    AccelBuild &ab = const_cast<AccelBuild &>(ab_in); // (allVectors is shared with the reader; the writer does not modify)
    Writer w;
    Scalars sc;
    std::memset(&sc, 0, sizeof sc);
    sc.root_link = ab.root_link, sc.n_leaves = ab.n_leaves, sc.ref_depth = ab.ref_depth;
    sc.root_is_reference_leaf = ab.root_is_reference_leaf ? 1 : 0, sc.n_sliver = ab.n_sliver, sc.n_needle = ab.n_needle;
    sc.wide_root = ab.wide_root, sc.wide_depth = ab.wide_depth, sc.miss_rank = ab.miss_rank, sc.scene_scale = ab.scene_scale;
    sc.sah_ref = ab.sah_ref, sc.sah_wide = ab.sah_wide;
    w.buf.append(reinterpret_cast<const char *>(&sc), sizeof sc);
    allVectors(w, ab);
    Header h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.magic, kMagic, 8);
    h.version = kFormatVersion, h.use_wide = use_wide ? 1u : 0u, h.key = key;
    h.payload_bytes = w.buf.size();
    h.checksum = fnv(w.buf.data(), w.buf.size(), 1469598103934665603ull);
    // (the process id in the name: several processes may create the same scene with the same cache path at once — each
    // writes its own file and the renames are atomic, so the survivor is one complete file)
    const std::string tmp = std::string(path) + ".tmp." + std::to_string((long long)getpid());
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f)
        return "cannot open " + tmp;
    const bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(w.buf.data(), 1, w.buf.size(), f) == w.buf.size();
    if (std::fclose(f) != 0 || !ok || std::rename(tmp.c_str(), path) != 0)
    {
        std::remove(tmp.c_str());
        return std::string("cannot write ") + path;
    }
    return "";
}

bool loadLayout(const char *path, uint64_t key, AccelBuild &ab, bool &use_wide, std::string &why)
{
    FILE *f = std::fopen(path, "rb");
    if (!f)
    {
        why = "no such file";
        return false;
    }
    Header h;
    std::string payload;
    why.clear();
    if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, kMagic, 8) != 0)
        why = "not a layout cache";
    else if (h.version != kFormatVersion)
        why = "other format version";
    else if (h.key != key)
        why = "made for another scene or other layout switches";
    else
    {
        std::fseek(f, 0, SEEK_END);
        const long size = std::ftell(f);
        if (size < 0 || (uint64_t)size != sizeof h + h.payload_bytes || h.payload_bytes < sizeof(Scalars))
            why = "truncated or oversized file";
        else
        {
            std::fseek(f, (long)sizeof h, SEEK_SET);
            payload.resize((size_t)h.payload_bytes);
            if (std::fread(&payload[0], 1, payload.size(), f) != payload.size())
                why = "read error";
            else if (fnv(payload.data(), payload.size(), 1469598103934665603ull) != h.checksum)
                why = "checksum mismatch";
        }
    }
    std::fclose(f);
    if (!why.empty())
        return false;
    Scalars sc;
    std::memcpy(&sc, payload.data(), sizeof sc);
    AccelBuild fresh;
    Reader r{payload.data() + sizeof sc, payload.data() + payload.size()};
    allVectors(r, fresh);
    if (!r.ok || r.p != r.end)
    {
        why = "malformed payload";
        return false;
    }
    fresh.root_link = sc.root_link, fresh.n_leaves = sc.n_leaves, fresh.ref_depth = sc.ref_depth;
    fresh.root_is_reference_leaf = sc.root_is_reference_leaf != 0, fresh.n_sliver = sc.n_sliver, fresh.n_needle = sc.n_needle;
    fresh.wide_root = sc.wide_root, fresh.wide_depth = sc.wide_depth, fresh.miss_rank = sc.miss_rank;
    fresh.scene_scale = sc.scene_scale, fresh.sah_ref = sc.sah_ref, fresh.sah_wide = sc.sah_wide;
    ab = std::move(fresh);
    use_wide = h.use_wide != 0;
    return true;
}
} // namespace trt
