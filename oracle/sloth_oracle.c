/*
 * sloth_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of rust-sloth's per-frame raster
 * path, used only as the checker for the CUDA path (tests/, smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).  Nothing in the product
 * library (rust-sloth_b200/) links, imports or falls back to this file.
 *
 * PARITY UNPINNED: the reference (Rust) cannot be compiled in this image (no
 * rustc/cargo, no vendored crates), and its own tests hold only three
 * known-answer vectors (src/geometry.rs:196-234), none of which exercises the
 * rasteriser.  The arithmetic below that lives in un-vendored crates
 * (nalgebra 0.22.1 + simba 0.2.1, Cargo.lock) is restated from their published
 * source: gemm -> gemv -> axcpy accumulation order, the 4-lane dot pairing in
 * norm_squared, Unit::new_normalize = divide by norm.  The pin available here
 * is: those three vectors + an independent numpy restatement
 * (oracle/restate_np.py) that must agree with this file bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (SSE2 scalar f32, no x87,
 * no FMA contraction), see oracle/Makefile.
 *
 * Every function cites the reference lines it follows (paths are relative to
 * the reference checkout).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* counters[]: 0 candidates (rasterizer.rs:70), 1 covered fragments (:75),
 *             2 z-buffer writes (:81-82), 3 newline stamps (:89-91)        */
enum { C_CAND = 0, C_COV = 1, C_ZW = 2, C_STAMP = 3, C_N = 4 };

/* default_shader thresholds and ramp, src/rasterizer.rs:5-27 */
static const float k_default_thr[9] = {0.20f, 0.30f, 0.40f, 0.50f, 0.60f,
                                       0.70f, 0.80f, 0.90f, 1.0f};
static const char k_default_glyph[10] = {'.', ':', '-', '=', '+',
                                         '*', '#', '%', '@', ' '};

/* Rust `f32 as usize`: saturating, NaN -> 0 (used at rasterizer.rs:59-66). */
static size_t sat_usize(float v)
{
    if (!(v > 0.0f)) return 0; /* negative, zero, NaN */
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}

/* Column-major 4x4 (nalgebra storage): element (row r, col c) = m[c*4 + r]. */
#define M_(m, r, c) ((m)[(c) * 4 + (r)])

/*
 * Matrix4 * Matrix4 as nalgebra 0.22.1 does it for statically sized matrices:
 * gemm loops over output columns, gemv accumulates column by column with
 * axcpy, i.e. C[i][j] = ((A[i][0]*B[0][j] + A[i][1]*B[1][j]) + A[i][2]*B[2][j])
 * + A[i][3]*B[3][j], every product and sum rounded to f32.
 * Call site: src/rasterizer.rs:57 (context.utransform * transform).
 */
ORACLE_API void oracle_mat4_mul(const float *A, const float *B, float *C)
{
    float out[16];
    for (int j = 0; j < 4; ++j) {
        for (int i = 0; i < 4; ++i) {
            float acc = M_(A, i, 0) * M_(B, 0, j);
            for (int k = 1; k < 4; ++k) acc = acc + M_(A, i, k) * M_(B, k, j);
            out[j * 4 + i] = acc;
        }
    }
    memcpy(C, out, sizeof out);
}

/* Matrix4 * Vector4, same gemv order.  Call site: src/geometry.rs:44-46. */
static void mat4_mul_vec(const float *M, const float v[4], float out[4])
{
    float r[4];
    for (int i = 0; i < 4; ++i) {
        float acc = M_(M, i, 0) * v[0];
        for (int k = 1; k < 4; ++k) acc = acc + M_(M, i, k) * v[k];
        r[i] = acc;
    }
    memcpy(out, r, sizeof r);
}

/*
 * Context::update, src/context.rs:93-141, image-mode branch: the terminal
 * size is (width as u16, height as u16); `scale0` is the max over meshes of
 * max(bbox.max.x, .y, .z) folded from 0.0 (:106-113, done by the caller).
 * Returns 0 and leaves `out` untouched when the u16 size is (0,0), because
 * the reference then skips the block (old_size == terminal_size, :104).
 */
ORACLE_API int oracle_utransform(uint32_t W, uint32_t H, float scale0, float *out)
{
    uint16_t w16 = (uint16_t)W, h16 = (uint16_t)H;
    if (w16 == 0 && h16 == 0) return 0;
    float fw = (float)w16, fh = (float)h16;
    float scale = fminf(fh, fw / 2.0f) / scale0 / 2.0f; /* :114 */
    memset(out, 0, 16 * sizeof(float));
    M_(out, 0, 0) = scale;
    M_(out, 0, 3) = fw / 4.0f;
    M_(out, 1, 1) = -scale;
    M_(out, 1, 3) = fh / 2.0f;
    M_(out, 2, 2) = scale;
    M_(out, 3, 3) = 1.0f;
    return 1;
}

/*
 * Rotation3::from_euler_angles(roll, pitch, yaw).to_homogeneous()
 * (nalgebra 0.22.1), call site src/main.rs:76-77.  sin/cos are the platform
 * libm's sinf/cosf (Rust f32::sin_cos).  Products are left-associated.
 */
ORACLE_API void oracle_rotation(float roll, float pitch, float yaw, float *out)
{
    float sr = sinf(roll), cr = cosf(roll);
    float sp = sinf(pitch), cp = cosf(pitch);
    float sy = sinf(yaw), cy = cosf(yaw);
    memset(out, 0, 16 * sizeof(float));
    M_(out, 0, 0) = cy * cp;
    M_(out, 0, 1) = cy * sp * sr - sy * cr;
    M_(out, 0, 2) = cy * sp * cr + sy * sr;
    M_(out, 1, 0) = sy * cp;
    M_(out, 1, 1) = sy * sp * sr + cy * cr;
    M_(out, 1, 2) = sy * sp * cr - cy * sr;
    M_(out, 2, 0) = -sp;
    M_(out, 2, 1) = cp * sr;
    M_(out, 2, 2) = cp * cr;
    M_(out, 3, 3) = 1.0f;
}

/*
 * The -j turntable angle sequence, src/main.rs:55-58,92-106 with
 * src/inputs.rs:131-149: pitch_0 = y + PI (f32), step = (2*PI)*(1/N) in f32,
 * pitch accumulates in f32; the loop stops after the frame where the
 * advanced pitch exceeds 9.42477 or N-1 frames have been counted.
 * Writes the pitch used by each rendered frame; returns the frame count.
 */
ORACLE_API size_t oracle_turntable(float y_arg, uint32_t n_frames, float *pitches, size_t cap)
{
    const float pi = 3.14159265358979323846f; /* std::f32::consts::PI */
    float pitch = y_arg;
    pitch += pi;                                           /* inputs.rs:148 */
    float step = (2.0f * pi) * (1.0f / (float)n_frames);   /* main.rs:57 */
    size_t count = 0;
    uint32_t frame_count = 0;
    for (;;) {
        if (count < cap) pitches[count] = pitch;
        count++;
        pitch += step;                                     /* main.rs:92-96 */
        /* main.rs:99; `webify_todo_frames - 1 == webify_frame_count` is i32
         * arithmetic in the reference (parse() infers i32). */
        if (pitch > 9.42477f || (int64_t)n_frames - 1 == (int64_t)frame_count) break;
        frame_count++;
    }
    return count;
}

/* orient, src/rasterizer.rs:30-32 */
static float orient(const float a[4], const float b[4], const float c[4])
{
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}

/*
 * Triangle::normal().z, src/geometry.rs:49-56: cross of (v2-v1),(v3-v1) in a
 * Vector4 with w = 0, normalised by Unit::new_normalize = each component
 * divided by norm(); norm_squared uses nalgebra's 4-lane dot, which sums
 * lanes as (l0 + l2) + (l1 + l3).
 */
static float normal_z(const float v1[4], const float v2[4], const float v3[4])
{
    float a[4], b[4];
    for (int i = 0; i < 4; ++i) {
        a[i] = v2[i] - v1[i];
        b[i] = v3[i] - v1[i];
    }
    float x = (a[1] * b[2]) - (a[2] * b[1]);
    float y = (a[2] * b[0]) - (a[0] * b[2]);
    float z = (a[0] * b[1]) - (a[1] * b[0]);
    float w = 0.0f;
    float l0 = x * x, l1 = y * y, l2 = z * z, l3 = w * w;
    l0 += l2;
    l1 += l3;
    float n2 = 0.0f;
    n2 += l0 + l1;
    float n = sqrtf(n2);
    return z / n;
}

/* the `shader: F` closure; default = default_shader, src/rasterizer.rs:5-27 */
static char shade_glyph(float shade, const float *thr, const char *glyph)
{
    for (int i = 0; i < 9; ++i)
        if (shade <= thr[i]) return glyph[i];
    return glyph[9];
}

static inline uint32_t pack_cell(char ch, uint8_t r, uint8_t g, uint8_t b)
{
    return (uint32_t)(uint8_t)ch | ((uint32_t)r << 8) | ((uint32_t)g << 16) | ((uint32_t)b << 24);
}

/*
 * One frame: Context::update + Context::clear + draw_mesh over the whole
 * triangle soup (src/main.rs:78-83).
 *
 *  xyz      N*9 floats (v1.xyz, v2.xyz, v3.xyz per triangle, draw order);
 *           w = 1.0 is implied (src/geometry.rs:85-101,158-171)
 *  rgb      N*3 bytes, Triangle.color
 *  scale0   max over meshes of max(bbox.max.xyz) folded from 0 (context.rs:106-113)
 *  image    Context.image (adds H cells and the '\n' stamps)
 *  rot      column-major 4x4, the `transform` argument of draw_mesh
 *  mode     0 = the reference scan domain, candidate by candidate
 *           1 = same results, but each row stops at the first candidate where
 *               an edge whose y-delta is >= 0 fails (that edge's value can
 *               only decrease further right, so nothing beyond can pass);
 *               only applied to triangles whose coordinates are finite and
 *               below 2^40, otherwise falls back to mode 0.  Used to check
 *               sizes where mode 0 takes minutes; validated against mode 0.
 *  tri_first/tri_step  process triangles tri_first, tri_first+tri_step, ...
 *           (bounded samples for the CPU baseline); 0/1 = all
 *  thr/glyph  shader table or NULL for default_shader
 *  cells    W*H (+H if image) packed cells: glyph | r<<8 | g<<16 | b<<24
 *  zbuf     W*H floats
 */
ORACLE_API int oracle_render(const float *xyz, const uint8_t *rgb, size_t n_tri, float scale0,
                             uint32_t W32, uint32_t H32, int image, const float *rot, int mode,
                             size_t tri_first, size_t tri_step, const float *thr,
                             const char *glyph, uint32_t *cells, float *zbuf, uint64_t *counters)
{
    const size_t W = W32, H = H32;
    uint64_t cnt[C_N] = {0, 0, 0, 0};
    if (!thr) thr = k_default_thr;
    if (!glyph) glyph = k_default_glyph;
    if (tri_step == 0) tri_step = 1;

    /* Context::blank + update (context.rs:22-34,93-141) */
    float utransform[16];
    memset(utransform, 0, sizeof utransform);
    M_(utransform, 0, 0) = M_(utransform, 1, 1) = M_(utransform, 2, 2) = M_(utransform, 3, 3) = 1.0f;
    oracle_utransform(W32, H32, scale0, utransform);

    /* Context::clear (context.rs:35-45) */
    const size_t n_cells = W * H + (image ? H : 0);
    const uint32_t blank = pack_cell(' ', 0, 0, 0);
    for (size_t i = 0; i < n_cells; ++i) cells[i] = blank;
    for (size_t i = 0; i < W * H; ++i) zbuf[i] = 3.40282347e+38f; /* f32::MAX */

    for (size_t t = tri_first; t < n_tri; t += tri_step) {
        /* draw_triangle, rasterizer.rs:48-93 */
        float M[16];
        oracle_mat4_mul(utransform, rot, M); /* :57, recomputed per triangle */
        float v1[4] = {xyz[t * 9 + 0], xyz[t * 9 + 1], xyz[t * 9 + 2], 1.0f};
        float v2[4] = {xyz[t * 9 + 3], xyz[t * 9 + 4], xyz[t * 9 + 5], 1.0f};
        float v3[4] = {xyz[t * 9 + 6], xyz[t * 9 + 7], xyz[t * 9 + 8], 1.0f};
        mat4_mul_vec(M, v1, v1); /* geometry.rs:43-48 */
        mat4_mul_vec(M, v2, v2);
        mat4_mul_vec(M, v3, v3);
        const uint8_t cr = rgb[t * 3 + 0], cg = rgb[t * 3 + 1], cb = rgb[t * 3 + 2];

        /* aabb, geometry.rs:37-42 (f32::min/max ignore NaN like fminf/fmaxf) */
        float mn0 = fminf(v1[0], fminf(v2[0], v3[0])), mn1 = fminf(v1[1], fminf(v2[1], v3[1]));
        float mx0 = fmaxf(v1[0], fmaxf(v2[0], v3[0])), mx1 = fmaxf(v1[1], fmaxf(v2[1], v3[1]));
        /* rasterizer.rs:59-66 */
        size_t minx = sat_usize(ceilf(fmaxf(mn0, 1.0f)));
        size_t miny = sat_usize(ceilf(fmaxf(mn1, 1.0f)));
        size_t maxx = sat_usize(ceilf(fminf(mx0 * 2.0f, (float)(W - 1))));
        size_t maxy = sat_usize(ceilf(fminf(mx1, (float)(H - 1))));
        float a = 1.0f / orient(v1, v2, v3); /* :67 */

        int fast = 0;
        if (mode == 1) {
            fast = 1;
            const float lim = 1099511627776.0f; /* 2^40 */
            for (int i = 0; i < 3; ++i)
                if (!(fabsf(v1[i]) <= lim) || !(fabsf(v2[i]) <= lim) || !(fabsf(v3[i]) <= lim))
                    fast = 0;
        }
        /* per-edge y deltas, only used by the mode-1 early break */
        const float dy0 = v3[1] - v2[1], dy1 = v1[1] - v3[1], dy2 = v2[1] - v1[1];

        for (size_t y = miny; y < maxy; ++y) {       /* :69 */
            for (size_t x = minx; x < maxx; ++x) {   /* :70 */
                float p[4] = {(float)x, (float)y, 0.0f, 0.0f};
                float w0 = orient(v2, v3, p);
                float w1 = orient(v3, v1, p);
                float w2 = orient(v1, v2, p);
                cnt[C_CAND]++;
                if (w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f) {     /* :75 */
                    cnt[C_COV]++;
                    float pixel_shade = normal_z(v1, v2, v3) * a * (w0 + w1 + w2);      /* :76 */
                    float z = v1[2] + a * (w1 * (v2[2] - v1[2]) + w2 * (v3[2] - v1[2])); /* :77-79 */
                    size_t id = y * W + x * 2;                                           /* :80 */
                    if (z < zbuf[id]) {                                                  /* :81 */
                        cnt[C_ZW]++;
                        zbuf[id] = z;
                        uint32_t px = pack_cell(shade_glyph(pixel_shade, thr, glyph), cr, cg, cb);
                        cells[id] = px;
                        cells[id + 1] = px;
                    }
                } else if (fast) {
                    if ((w0 < 0.0f && !(dy0 < 0.0f)) || (w1 < 0.0f && !(dy1 < 0.0f)) ||
                        (w2 < 0.0f && !(dy2 < 0.0f)))
                        break;
                }
            }
            if (image) {                                          /* :89-91 */
                cells[y * W + 1] = pack_cell('\n', 0, 0, 0);
                cnt[C_STAMP]++;
            }
        }
    }
    if (counters) memcpy(counters, cnt, sizeof cnt);
    return 0;
}

/* Known-answer helpers for src/geometry.rs:196-234 (test_aabb, test_transform,
 * test_normal): expose aabb / mul / normal on one triangle. */
ORACLE_API void oracle_triangle_aabb(const float *v /*12: three Vector4*/, float *mn, float *mx)
{
    for (int i = 0; i < 4; ++i) {
        mn[i] = fminf(v[i], fminf(v[4 + i], v[8 + i]));
        mx[i] = fmaxf(v[i], fmaxf(v[4 + i], v[8 + i]));
    }
}

ORACLE_API void oracle_triangle_mul(const float *M, float *v /*12, in place*/)
{
    mat4_mul_vec(M, v, v);
    mat4_mul_vec(M, v + 4, v + 4);
    mat4_mul_vec(M, v + 8, v + 8);
}

ORACLE_API void oracle_triangle_normal(const float *v /*12*/, float *out /*4*/)
{
    float a[4], b[4];
    for (int i = 0; i < 4; ++i) {
        a[i] = v[4 + i] - v[i];
        b[i] = v[8 + i] - v[i];
    }
    float n[4];
    n[0] = (a[1] * b[2]) - (a[2] * b[1]);
    n[1] = (a[2] * b[0]) - (a[0] * b[2]);
    n[2] = (a[0] * b[1]) - (a[1] * b[0]);
    n[3] = 0.0f;
    float l0 = n[0] * n[0], l1 = n[1] * n[1], l2 = n[2] * n[2], l3 = n[3] * n[3];
    l0 += l2;
    l1 += l3;
    float nn = sqrtf(0.0f + (l0 + l1));
    for (int i = 0; i < 4; ++i) out[i] = n[i] / nn;
}

/* Unit::new_normalize on a plain Vector4 (right-hand side of test_normal). */
ORACLE_API void oracle_vec4_normalize(const float *v, float *out)
{
    float l0 = v[0] * v[0], l1 = v[1] * v[1], l2 = v[2] * v[2], l3 = v[3] * v[3];
    l0 += l2;
    l1 += l3;
    float nn = sqrtf(0.0f + (l0 + l1));
    for (int i = 0; i < 4; ++i) out[i] = v[i] / nn;
}
