// loader_api.inl -- C entry points of the device loader (kernels: loader.cuh).  Part of capi.cu's translation
// unit: uses sloth_ctx, CU(), fail(), alloc_scene(), finish_scene().
//
// Host work here is limited to moving the file's bytes (read() into two page-locked staging buffers that the copy
// engine drains alternately) and to what is O(1) in the model size: the ASCII / binary STL test (stl_io looks at
// the first bytes), fetching the (few) mtllib lines the device reports and parsing those .mtl files
// (host/mesh_io.cpp), error text.
#include <chrono>
#include <string>
#include <vector>

#include "../host/mesh_io.hpp"

struct LoaderState {
    struct Segment {
        float* xyz = nullptr;     // n_tri * 9, device
        uint8_t* rgb = nullptr;   // n_tri * 3, device
        size_t n_tri = 0;
    };
    std::vector<Segment> segments;
    cudaStream_t stream = nullptr;
    // file -> device staging: two page-locked buffers, filled by read() while the other one is being copied
    static constexpr size_t STAGE_BYTES = 8u << 20;
    unsigned char* stage[2] = {nullptr, nullptr};
    cudaEvent_t stage_free[2] = {nullptr, nullptr};
    void clear()
    {
        for (auto& s : segments) { cudaFreeAsync(s.xyz, stream); cudaFreeAsync(s.rgb, stream); }
        segments.clear();
    }
    ~LoaderState()
    {
        clear();
        for (int i = 0; i < 2; ++i) {
            if (stage[i]) cudaFreeHost(stage[i]);
            if (stage_free[i]) cudaEventDestroy(stage_free[i]);
        }
    }
};

namespace {

// Device allocation from the stream-ordered pool, released (in stream order) on scope exit.  One model load
// allocates and drops a dozen temporaries of up to hundreds of MB; the pool hands the same memory on from one to
// the next instead of mapping it afresh, and freeing never stalls the device.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t n) { return cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st); }
    T* release() { T* q = p; p = nullptr; return q; }
};

// keep freed blocks in the pool while a load is in progress; give everything back afterwards
void ld_pool_hold(sloth_ctx* c, bool hold)
{
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) != cudaSuccess) { cudaGetLastError(); return; }
    unsigned long long thr = hold ? ~0ull : 0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    if (!hold) {
        cudaStreamSynchronize(c->stream);
        cudaMemPoolTrimTo(pool, 0);
    }
}

struct LoaderOut {   // small device block read back after the counting passes
    ld::ObjTotals tot;
    unsigned long long err_parse, err_export;
    uint32_t mtllib_count, n_newlines;
    uint32_t mtllib_lines[ld::LD_MAX_MTLLIB];
};

const char* ld_message(uint32_t code)
{
    switch (code) {
    case ld::LD_POSITION: return "position parse error";
    case ld::LD_FACE: return "face parse error";
    case ld::LD_MATERIAL: return "material parse error";
    case ld::LD_UNSUPPORTED: return "token outside the device parser's grammar";
    case ld::LD_NO_MATERIAL:
        return "model has no material although the material list is non-empty (the reference panics on "
               "material_id.unwrap(), geometry.rs:110)";
    case ld::LD_MISSING_VERTEX: return "face references a missing vertex";
    case ld::LD_VCOL_RANGE: return "vertex colour index out of range";
    case ld::LD_STL_VERTEX: return "stl_io couldnt parse STL: bad vertex";
    case ld::LD_TOO_MANY_MTLLIB: return "more mtllib statements than the device loader tracks";
    default: return "unknown loader error";
    }
}

int ld_fail(unsigned long long word)
{
    const uint32_t code = (uint32_t)(word & 0xFFu);
    const unsigned long long line = (word >> 8) + 1ull;
    const int rc = (code == ld::LD_UNSUPPORTED || code == ld::LD_TOO_MANY_MTLLIB) ? SLOTH_E_UNSUPPORTED : SLOTH_E_PARSE;
    return fail(rc, "%s (line %llu)", ld_message(code), line);
}

// newline positions -> line table for text that is already on the device (enqueued on c->stream)
int ld_lines(sloth_ctx* c, const unsigned char* d_text, size_t len, DevBuf<uint32_t>& d_line_start, LoaderOut* d_out, uint32_t& n_lines)
{
    const uint32_t n_tiles = (uint32_t)std::max<size_t>(1, (len + ld::LD_TILE - 1) / ld::LD_TILE);
    DevBuf<uint32_t> d_tiles(c->stream);
    CU(d_tiles.alloc(n_tiles));
    ld::k_count_newlines<<<n_tiles, 256, 0, c->stream>>>(d_text, (uint32_t)len, d_tiles.p);
    ld::k_scan_tiles<<<1, 1024, 0, c->stream>>>(d_tiles.p, n_tiles, &d_out->n_newlines);
    uint32_t n_newlines = 0;
    CU(cudaMemcpyAsync(&n_newlines, &d_out->n_newlines, sizeof n_newlines, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    n_lines = n_newlines + 1u;
    CU(d_line_start.alloc((size_t)n_lines + 1));
    ld::k_line_starts<<<n_tiles, 256, 0, c->stream>>>(d_text, (uint32_t)len, d_tiles.p, n_newlines, d_line_start.p);
    c->launches += 3;
    CU(cudaGetLastError());
    return SLOTH_OK;
}

// host bytes -> device (pageable or pinned source)
int ld_upload(sloth_ctx* c, const void* bytes, size_t len, DevBuf<unsigned char>& d_text)
{
    CU(d_text.alloc(len + 64));
    if (len) CU(cudaMemcpyAsync(d_text.p, bytes, len, cudaMemcpyHostToDevice, c->stream));
    return SLOTH_OK;
}

int ld_scan(sloth_ctx* c, const uint4* d_rec, uint32_t n_lines, uint4* d_pre, LoaderOut* d_out)
{
    const uint32_t n_blocks = (n_lines + ld::SCAN_PER_BLOCK - 1) / ld::SCAN_PER_BLOCK;
    DevBuf<uint4> d_block(c->stream);
    CU(d_block.alloc(n_blocks));
    ld::k_scan_reduce<<<n_blocks, ld::SCAN_THREADS, 0, c->stream>>>(d_rec, n_lines, d_block.p);
    ld::k_scan_blocks<<<1, ld::SCAN_THREADS, 0, c->stream>>>(d_block.p, n_blocks, &d_out->tot);
    ld::k_scan_apply<<<n_blocks, ld::SCAN_THREADS, 0, c->stream>>>(d_rec, n_lines, d_block.p, d_pre);
    c->launches += 3;
    CU(cudaGetLastError());
    return SLOTH_OK;
}

int ld_init_out(sloth_ctx* c, DevBuf<LoaderOut>& d_out)
{
    CU(d_out.alloc(1));
    LoaderOut init;
    std::memset(&init, 0, sizeof init);
    init.err_parse = init.err_export = ~0ull;
    CU(cudaMemcpyAsync(d_out.p, &init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));   // `init` leaves scope
    return SLOTH_OK;
}

bool host_is_ws(unsigned char ch) { return ch == ' ' || (ch >= 9 && ch <= 13); }

// OBJ text that is already on the device (len bytes at d_text.p)
int loader_add_obj(sloth_ctx* c, DevBuf<unsigned char>& d_text, size_t len, const char* mtl_dir)
{
    DevBuf<uint32_t> d_line_start(c->stream);
    DevBuf<LoaderOut> d_out(c->stream);
    int rc = ld_init_out(c, d_out);
    if (rc) return rc;
    uint32_t n_lines = 0;
    rc = ld_lines(c, d_text.p, len, d_line_start, d_out.p, n_lines);
    if (rc) return rc;

    DevBuf<uint4> d_rec(c->stream), d_pre(c->stream);
    CU(d_rec.alloc(n_lines));
    CU(d_pre.alloc(n_lines));
    const uint32_t line_blocks = (n_lines + 255u) / 256u;
    ld::k_obj_classify<<<line_blocks, 256, 0, c->stream>>>(d_text.p, d_line_start.p, n_lines, d_rec.p, d_out.p->mtllib_lines,
                                                            &d_out.p->mtllib_count, &d_out.p->err_parse);
    c->launches += 1;
    LoaderOut out;
    CU(cudaMemcpyAsync(&out, d_out.p, sizeof out, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));

    // mtllib statements (tobj loads each file when it reads the statement; a usemtl only sees what is loaded by then)
    std::vector<ld::Material> mats;
    std::string names;
    bool mtl_failed = false;
    std::string mtl_err;
    if (out.mtllib_count && out.mtllib_count <= ld::LD_MAX_MTLLIB) {
        std::sort(out.mtllib_lines, out.mtllib_lines + out.mtllib_count);
        for (uint32_t k = 0; k < out.mtllib_count; ++k) {
            uint32_t span[2] = {0u, 0u};   // start of the line and of the next one
            CU(cudaMemcpyAsync(span, d_line_start.p + out.mtllib_lines[k], sizeof span, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            const uint32_t off = span[0];
            std::string line(span[1] - 1u - off, '\0');
            if (!line.empty()) CU(cudaMemcpyAsync(&line[0], d_text.p + off, line.size(), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            // second token of the line
            size_t i = 0;
            auto skip_ws = [&] { while (i < line.size() && host_is_ws((unsigned char)line[i])) ++i; };
            auto skip_tok = [&] { while (i < line.size() && !host_is_ws((unsigned char)line[i])) ++i; };
            skip_ws(); skip_tok(); skip_ws();
            const size_t b = i;
            skip_tok();
            const std::string file = line.substr(b, i - b);
            std::vector<sloth::MtlMaterial> loaded;
            std::string e;
            if (!sloth::load_mtl_file(std::string(mtl_dir ? mtl_dir : "") + file, loaded, e)) {
                mtl_failed = true;
                mtl_err = e;
                continue;
            }
            for (const auto& m : loaded) {
                ld::Material dm;
                dm.name_off = (uint32_t)names.size();
                dm.name_len = (uint32_t)m.name.size();
                dm.defined_at = off;
                dm.rgb = (uint32_t)sloth::f32_as_u8(m.diffuse[0] * 255.0f) | (uint32_t)sloth::f32_as_u8(m.diffuse[1] * 255.0f) << 8 |
                         (uint32_t)sloth::f32_as_u8(m.diffuse[2] * 255.0f) << 16;
                names += m.name;
                mats.push_back(dm);
            }
        }
    }
    const uint32_t n_mats = (uint32_t)mats.size();
    DevBuf<ld::Material> d_mats(c->stream);
    DevBuf<unsigned char> d_names(c->stream);
    CU(d_mats.alloc(n_mats));
    CU(d_names.alloc(names.size()));
    if (n_mats) {
        CU(cudaMemcpyAsync(d_mats.p, mats.data(), n_mats * sizeof(ld::Material), cudaMemcpyHostToDevice, c->stream));
        if (!names.empty()) CU(cudaMemcpyAsync(d_names.p, names.data(), names.size(), cudaMemcpyHostToDevice, c->stream));
        ld::k_obj_usemtl<<<line_blocks, 256, 0, c->stream>>>(d_text.p, d_line_start.p, n_lines, d_rec.p, d_mats.p, n_mats, d_names.p);
        c->launches += 1;
    }
    rc = ld_scan(c, d_rec.p, n_lines, d_pre.p, d_out.p);
    if (rc) return rc;
    CU(cudaMemcpyAsync(&out, d_out.p, sizeof out, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (out.err_parse != ~0ull) return ld_fail(out.err_parse);
    // inputs.rs:112; tobj 3.2.2 returns Ok(materials) whenever the final material list is non-empty
    if (mtl_failed && n_mats == 0) return fail(SLOTH_E_IO, "Expected to have materials. (%s)", mtl_err.c_str());
    const size_t n_tri = out.tot.n_tris;
    if (n_tri > MAX_TRIS) return fail(SLOTH_E_TOO_LARGE, "%zu triangles; the depth key holds a 27-bit index (max %u)", n_tri, MAX_TRIS);

    DevBuf<float> d_pos(c->stream), d_vcol(c->stream), d_xyz(c->stream);
    DevBuf<uint8_t> d_rgb(c->stream);
    CU(d_pos.alloc((size_t)out.tot.n_vertices * 3));
    CU(d_vcol.alloc(out.tot.n_vcol));
    CU(d_xyz.alloc(n_tri * 9));
    CU(d_rgb.alloc(n_tri * 3));
    ld::k_obj_vertices<<<line_blocks, 256, 0, c->stream>>>(d_text.p, d_line_start.p, n_lines, d_rec.p, d_pre.p, d_pos.p, d_vcol.p);
    ld::k_obj_faces<<<line_blocks, 256, 0, c->stream>>>(d_text.p, d_line_start.p, n_lines, d_rec.p, d_pre.p, d_pos.p, d_vcol.p, out.tot,
                                                         d_mats.p, n_mats, d_xyz.p, d_rgb.p, &d_out.p->err_export);
    c->launches += 2;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&out.err_export, &d_out.p->err_export, sizeof out.err_export, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (out.err_export != ~0ull) return ld_fail(out.err_export);
    // (material_id.unwrap() sits inside the reference's per-triangle loop, geometry.rs:109-110: k_obj_faces reports
    // LD_NO_MATERIAL for every face that emits a triangle without a material; a trailing model without faces is fine)
    LoaderState::Segment seg;
    seg.xyz = d_xyz.release();
    seg.rgb = d_rgb.release();
    seg.n_tri = n_tri;
    c->loader->segments.push_back(seg);
    return SLOTH_OK;
}

// STL bytes that are already on the device; `head` = the first min(len, 4096) bytes on the host
int loader_add_stl(sloth_ctx* c, DevBuf<unsigned char>& d_text, size_t len, const unsigned char* head)
{
    // stl_io::create_stl_reader: AsciiStlReader::probe decides (first line valid UTF-8 and starting with "solid ";
    // host/mesh_io.cpp).  A first line longer than the 4 KB the host sees is judged on those 4 KB.
    const size_t head_len = std::min<size_t>(len, 4096);
    const bool ascii = sloth::stl_probe_ascii(head, head_len);
    DevBuf<float> d_xyz(c->stream);
    DevBuf<uint8_t> d_rgb(c->stream);
    size_t n_tri = 0;
    if (ascii) {
        DevBuf<uint32_t> d_line_start(c->stream);
        DevBuf<LoaderOut> d_out(c->stream);
        int rc = ld_init_out(c, d_out);
        if (rc) return rc;
        uint32_t n_lines = 0;
        rc = ld_lines(c, d_text.p, len, d_line_start, d_out.p, n_lines);
        if (rc) return rc;
        DevBuf<uint4> d_rec(c->stream), d_pre(c->stream);
        CU(d_rec.alloc(n_lines));
        CU(d_pre.alloc(n_lines));
        const uint32_t line_blocks = (n_lines + 255u) / 256u;
        ld::k_stl_ascii_classify<<<line_blocks, 256, 0, c->stream>>>(d_text.p, d_line_start.p, n_lines, d_rec.p, &d_out.p->err_parse);
        c->launches += 1;
        rc = ld_scan(c, d_rec.p, n_lines, d_pre.p, d_out.p);
        if (rc) return rc;
        LoaderOut out;
        CU(cudaMemcpyAsync(&out, d_out.p, sizeof out, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (out.err_parse != ~0ull) return ld_fail(out.err_parse);
        if (out.tot.n_vertices % 3u) return fail(SLOTH_E_PARSE, "stl_io couldnt parse STL: truncated facet");
        n_tri = out.tot.n_vertices / 3u;
        if (n_tri == 0 && len >= 84) {   // binary body behind a header that passes the probe (see host/mesh_io.cpp)
            uint32_t n = 0;
            std::memcpy(&n, head + 80, 4);
            if (n != 0 && len == 84 + (size_t)n * 50)
                return fail(SLOTH_E_PARSE, "stl_io couldnt parse STL: ASCII header (\"solid \") on a binary body");
        }
        if (n_tri > MAX_TRIS) return fail(SLOTH_E_TOO_LARGE, "%zu triangles; the depth key holds a 27-bit index (max %u)", n_tri, MAX_TRIS);
        CU(d_xyz.alloc(n_tri * 9));
        ld::k_stl_ascii_vertices<<<line_blocks, 256, 0, c->stream>>>(d_text.p, d_line_start.p, n_lines, d_rec.p, d_pre.p, d_xyz.p);
        c->launches += 1;
    } else {
        if (len < 84) return fail(SLOTH_E_PARSE, "stl_io couldnt parse STL: short binary header");
        uint32_t n = 0;
        std::memcpy(&n, head + 80, 4);
        if (len < 84 + (size_t)n * 50) return fail(SLOTH_E_PARSE, "stl_io couldnt parse STL: truncated binary body");
        n_tri = n;
        if (n_tri > MAX_TRIS) return fail(SLOTH_E_TOO_LARGE, "%zu triangles; the depth key holds a 27-bit index (max %u)", n_tri, MAX_TRIS);
        CU(d_xyz.alloc(n_tri * 9));
        if (n_tri) {
            ld::k_stl_binary<<<(unsigned)((n_tri * 9 + 255) / 256), 256, 0, c->stream>>>(d_text.p, (uint32_t)n_tri, d_xyz.p);
            c->launches += 1;
        }
    }
    CU(d_rgb.alloc(n_tri * 3));
    if (n_tri) {   // geometry.rs:161-162: every STL triangle is (0xFF, 0xFF, 0x00)
        ld::k_fill_rgb<<<(unsigned)((n_tri * 3 + 255) / 256), 256, 0, c->stream>>>(d_rgb.p, n_tri, 0x00FFFFu);
        c->launches += 1;
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    LoaderState::Segment seg;
    seg.xyz = d_xyz.release();
    seg.rgb = d_rgb.release();
    seg.n_tri = n_tri;
    c->loader->segments.push_back(seg);
    return SLOTH_OK;
}

int loader_commit(sloth_ctx* c, size_t* n_tri_out, float* scene_max_out)
{
    size_t total = 0;
    for (const auto& s : c->loader->segments) total += s.n_tri;
    if (total > MAX_TRIS) return fail(SLOTH_E_TOO_LARGE, "%zu triangles; the depth key holds a 27-bit index (max %u)", total, MAX_TRIS);
    int rc = alloc_scene(c, total);
    if (rc) return rc;
    DevBuf<uint32_t> d_stats(c->stream);
    CU(d_stats.alloc(2));
    CU(cudaMemsetAsync(d_stats.p, 0, 2 * sizeof(uint32_t), c->stream));
    size_t base = 0;
    for (const auto& s : c->loader->segments) {
        if (!s.n_tri) continue;
        const size_t nf = s.n_tri * 9;
        const unsigned blocks = (unsigned)std::min<size_t>((nf + 255) / 256, (size_t)c->sm_count * 16);
        ld::k_soup_scan<<<blocks, 256, 0, c->stream>>>(s.xyz, nf, d_stats.p);
        k_pack_scene<<<(unsigned)((s.n_tri + 255) / 256), 256, 0, c->stream>>>(s.xyz, s.rgb, (uint32_t)s.n_tri, c->sc_a + base, c->sc_b + base,
                                                                             c->sc_z3 + base, c->sc_rgb + base);
        c->launches += 2;
        base += s.n_tri;
    }
    rc = finish_scene(c, total);
    if (rc) return rc;
    uint32_t stats[2] = {0u, 0u};
    CU(cudaMemcpyAsync(stats, d_stats.p, sizeof stats, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float scene_max;
    std::memcpy(&scene_max, &stats[0], sizeof scene_max);
    c->scene_clean = stats[1] == 0u;
    c->n_tri = (uint32_t)total;
    c->scene_max = scene_max;
    c->have_scene = true;
    c->loader->clear();
    if (n_tri_out) *n_tri_out = total;
    if (scene_max_out) *scene_max_out = scene_max;
    return SLOTH_OK;
}

// file -> device through the two page-locked staging buffers: read() fills one while the copy engine drains the
// other.  `head` receives the first bytes (STL flavour test, binary STL count).
int ld_read_file(sloth_ctx* c, const std::string& path, DevBuf<unsigned char>& d_text, size_t* len_out, unsigned char head[4096])
{
    LoaderState* L = c->loader;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail(SLOTH_E_IO, "open failed");
    struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{f};
    if (std::fseek(f, 0, SEEK_END) != 0) return fail(SLOTH_E_IO, "open failed");
    const long sz = std::ftell(f);
    if (sz < 0 || std::fseek(f, 0, SEEK_SET) != 0) return fail(SLOTH_E_IO, "open failed");
    const size_t len = (size_t)sz;
    if (len >= 0xFFFFFF00ull) return fail(SLOTH_E_TOO_LARGE, "model file of %zu bytes; the device loader indexes bytes with 32 bits", len);
    for (int i = 0; i < 2; ++i) {
        if (!L->stage[i]) CU(cudaHostAlloc(&L->stage[i], LoaderState::STAGE_BYTES, cudaHostAllocDefault));
        if (!L->stage_free[i]) CU(cudaEventCreateWithFlags(&L->stage_free[i], cudaEventDisableTiming));
    }
    CU(d_text.alloc(len + 64));
    std::memset(head, 0, 4096);
    size_t off = 0;
    for (int k = 0; off < len; k ^= 1) {
        CU(cudaEventSynchronize(L->stage_free[k]));   // the copy that last used this buffer is done
        const size_t want = std::min(LoaderState::STAGE_BYTES, len - off);
        const size_t got = std::fread(L->stage[k], 1, want, f);
        if (got != want) return fail(SLOTH_E_IO, "read failed");
        if (off == 0) std::memcpy(head, L->stage[k], std::min<size_t>(got, 4096));
        CU(cudaMemcpyAsync(d_text.p + off, L->stage[k], got, cudaMemcpyHostToDevice, c->stream));
        CU(cudaEventRecord(L->stage_free[k], c->stream));
        off += got;
    }
    *len_out = len;
    return SLOTH_OK;
}

}  // namespace

extern "C" {

int sloth_loader_begin(sloth_ctx* c)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (!c->loader) c->loader = new (std::nothrow) LoaderState();
    if (!c->loader) return fail(SLOTH_E_ARG, "out of host memory");
    c->loader->stream = c->stream;
    c->loader->clear();
    ld_pool_hold(c, true);
    return SLOTH_OK;
}

int sloth_loader_add_obj(sloth_ctx* c, const char* text, size_t len, const char* mtl_dir)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (!c->loader) return fail(SLOTH_E_STATE, "sloth_loader_begin has not been called");
    if (len && !text) return fail(SLOTH_E_ARG, "text is null");
    if (len >= 0xFFFFFF00ull) return fail(SLOTH_E_TOO_LARGE, "model text of %zu bytes; the device loader indexes text with 32 bits", len);
    CU(cudaSetDevice(c->device));
    DevBuf<unsigned char> d_text(c->stream);
    const int rc = ld_upload(c, text, len, d_text);
    return rc ? rc : loader_add_obj(c, d_text, len, mtl_dir);
}

int sloth_loader_add_stl(sloth_ctx* c, const void* bytes, size_t len)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (!c->loader) return fail(SLOTH_E_STATE, "sloth_loader_begin has not been called");
    if (len && !bytes) return fail(SLOTH_E_ARG, "bytes is null");
    if (len >= 0xFFFFFF00ull) return fail(SLOTH_E_TOO_LARGE, "model file of %zu bytes; the device loader indexes bytes with 32 bits", len);
    CU(cudaSetDevice(c->device));
    DevBuf<unsigned char> d_text(c->stream);
    const int rc = ld_upload(c, bytes, len, d_text);
    return rc ? rc : loader_add_stl(c, d_text, len, static_cast<const unsigned char*>(bytes));
}

int sloth_loader_commit(sloth_ctx* c, size_t* n_tri_out, float* scene_max_out)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (!c->loader) return fail(SLOTH_E_STATE, "sloth_loader_begin has not been called");
    CU(cudaSetDevice(c->device));
    const int rc = loader_commit(c, n_tri_out, scene_max_out);
    ld_pool_hold(c, false);
    return rc;
}

// match_meshes, inputs.rs:95-129 (same splitting, extension rules and message format as host/mesh_io.cpp)
int sloth_scene_load(sloth_ctx* c, const char* models_arg, size_t* n_tri_out, float* scene_max_out)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (!models_arg) return fail(SLOTH_E_ARG, "models_arg is null");
    int rc = sloth_loader_begin(c);
    if (rc) return rc;
    const std::string arg(models_arg);
    using clk = std::chrono::steady_clock;
    auto ms_since = [](clk::time_point t) { return std::chrono::duration<float, std::milli>(clk::now() - t).count(); };
    c->load_ms[0] = c->load_ms[1] = c->load_ms[2] = 0.0f;
    size_t start = 0;
    for (;;) {
        const size_t sp = arg.find(' ', start);
        const std::string slice = arg.substr(start, sp == std::string::npos ? std::string::npos : sp - start);
        const size_t slash = slice.find_last_of('/');
        const std::string fname = slash == std::string::npos ? slice : slice.substr(slash + 1);
        const std::string dir = slash == std::string::npos ? std::string() : slice.substr(0, slash + 1);
        const size_t dot = fname.find_last_of('.');
        if (dot == std::string::npos || dot == 0) {
            c->loader->clear();
            ld_pool_hold(c, false);
            return fail(SLOTH_E_ARG, "filename: [%s] couldn't load, couldn't determine filename extension. ", slice.c_str());
        }
        std::string ext = fname.substr(dot + 1);
        std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char ch) { return std::tolower(ch); });
        const bool obj = ext == "obj", stl = ext == "stl";
        if (!obj && !stl) {
            c->loader->clear();
            ld_pool_hold(c, false);
            return fail(SLOTH_E_ARG, "filename: [%s] couldn't load, unknown filename extension. ", slice.c_str());
        }
        {
            DevBuf<unsigned char> d_text(c->stream);
            unsigned char head[4096];
            size_t len = 0;
            auto t0 = clk::now();
            rc = ld_read_file(c, slice, d_text, &len, head);
            c->load_ms[0] += ms_since(t0);
            if (rc == SLOTH_OK) {
                t0 = clk::now();
                rc = obj ? loader_add_obj(c, d_text, len, dir.c_str()) : loader_add_stl(c, d_text, len, head);
                c->load_ms[1] += ms_since(t0);
            }
        }
        if (rc) {
            c->loader->clear();
            ld_pool_hold(c, false);
            const std::string inner(g_err);
            const char* what = obj ? "tobj couldnt load/parse OBJ" : (rc == SLOTH_E_IO ? "STL load failed" : "stl_io couldnt parse STL");
            return fail(rc, "filename: [%s] couldn't load, %s. %s", slice.c_str(), what, inner.c_str());
        }
        if (sp == std::string::npos) break;
        start = sp + 1;
    }
    const auto t0 = clk::now();
    rc = sloth_loader_commit(c, n_tri_out, scene_max_out);
    c->load_ms[2] = ms_since(t0);
    return rc;
}

size_t sloth_scene_size(const sloth_ctx* c) { return (c && c->have_scene) ? c->n_tri : 0; }

int sloth_scene_get(sloth_ctx* c, float* xyz, uint8_t* rgb, float* scene_max_out)
{
    if (!c) return fail(SLOTH_E_ARG, "null context");
    if (!c->have_scene) return fail(SLOTH_E_STATE, "no scene is resident");
    CU(cudaSetDevice(c->device));
    const size_t n = c->n_tri;
    if (n && (xyz || rgb)) {
        DevBuf<float> d_xyz(c->stream);
        DevBuf<uint8_t> d_rgb(c->stream);
        CU(d_xyz.alloc(n * 9));
        CU(d_rgb.alloc(n * 3));
        ld::k_unpack_scene<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->sc_a, c->sc_b, c->sc_z3, c->sc_rgb, (uint32_t)n, d_xyz.p, d_rgb.p);
        c->launches += 1;
        CU(cudaGetLastError());
        if (xyz) CU(cudaMemcpyAsync(xyz, d_xyz.p, n * 9 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        if (rgb) CU(cudaMemcpyAsync(rgb, d_rgb.p, n * 3, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    if (scene_max_out) *scene_max_out = c->scene_max;
    return SLOTH_OK;
}

}  // extern "C"
