// sloth_main.cpp -- drop-in `sloth` command line over libsloth_b200.so (C++17 host).
//
// Mirrors the reference's host program around the raster path:
//   src/inputs.rs:9-85    clap definition: `sloth <input filename(s)> [-x/--yaw f] [-y/--pitch f]
//                         [-z/--roll f] [-b] [image -w W [-h H] [-j/--webify N] [-x -y -z -b]]`
//   src/main.rs:25-114    frame loop: interactive (raw mode, 500 fps cap, q / Ctrl-C), `image`
//                         single shot, `image -j N` JS-frame export
//   src/context.rs:50-92  flush: plain glyphs / ANSI truecolor per cell / <span> per cell -- serialised on
//                         the GPU (sloth_render_text*), the host only adds the frame prefixes/suffixes
// The per-frame group update + clear + draw_mesh (main.rs:78-83) is one sloth_render call.
// Quirks kept on purpose: only the first positional value is read and split on ' '
// (inputs.rs:97); in image mode the rotation flags are taken from the sub-command only
// (main.rs:43) and `-b` from the top level only (main.rs:34); `-h` is the height inside `image`.
#include <poll.h>
#include <signal.h>
#include <sys/ioctl.h>
#include <termios.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/sloth_b200.h"
#include "mesh_io.hpp"

namespace {

struct Matches {
    std::map<std::string, std::string> values;  // "x","y","z","width","height","frame count","input"
    bool no_color = false;
    bool present = false;
};

[[noreturn]] void usage_error(const std::string& msg)
{
    std::fprintf(stderr, "error: %s\n\nUSAGE:\n    sloth [FLAGS] [OPTIONS] <input filename(s)>... [SUBCOMMAND]\n\n"
                         "For more information try --help\n", msg.c_str());
    std::exit(1);
}

void print_help()
{
    std::puts("Sloth 0.1\nMitchell Hynes. <mshynes@mun.ca>\nA toy for rendering 3D objects in the command line\n\n"
              "USAGE:\n    sloth [FLAGS] [OPTIONS] <input filename(s)>... [SUBCOMMAND]\n\n"
              "FLAGS:\n    -b               Flags the rasterizer to render without color\n"
              "        --help       Prints help information\n    -V, --version    Prints version information\n\n"
              "OPTIONS:\n    -x, --yaw <x>      Sets the object's static X rotation (in radians)\n"
              "    -y, --pitch <y>    Sets the object's static Y rotation (in radians)\n"
              "    -z, --roll <z>     Sets the object's static Z rotation (in radians)\n\n"
              "ARGS:\n    <input filename(s)>...    Sets the input file to render\n\n"
              "SUBCOMMANDS:\n    image    Generates a colorless terminal output as lines of text\n"
              "             image -w <width> [-h <height>] [-j, --webify <frame count>] [-x -y -z -b]");
}

// clap 2 style parse of one level; returns the index of the sub-command token or argc.
int parse_level(int argc, char** argv, int i, bool sub, Matches& m)
{
    m.present = true;
    for (; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&](const char* name) {
            if (i + 1 >= argc) usage_error(std::string("The argument '") + name + "' requires a value but none was supplied");
            // clap 2 without allow_hyphen_values (inputs.rs sets none): a value that starts with '-' is taken for
            // the next flag, so `-x -1.5` is an error there, too
            if (argv[i + 1][0] == '-' && argv[i + 1][1] != '\0')
                usage_error(std::string("Found argument '") + argv[i + 1] + "' which wasn't expected, or isn't valid in this context");
            return std::string(argv[++i]);
        };
        if (a == "-x" || a == "--yaw") m.values["x"] = value("--yaw <x>");
        else if (a == "-y" || a == "--pitch") m.values["y"] = value("--pitch <y>");
        else if (a == "-z" || a == "--roll") m.values["z"] = value("--roll <z>");
        else if (a == "-b") m.no_color = true;
        else if (sub && a == "-w") m.values["width"] = value("-w <width>");
        else if (sub && a == "-h") m.values["height"] = value("-h <height>");
        else if (sub && (a == "-j" || a == "--webify")) m.values["frame count"] = value("--webify <frame count>");
        else if (!sub && (a == "--help" || a == "-h")) { print_help(); std::exit(0); }
        else if (!sub && (a == "-V" || a == "--version")) { std::puts("Sloth 0.1"); std::exit(0); }
        else if (!sub && a == "image") return i;
        else if (!a.empty() && a[0] == '-' && a.size() > 1 && !std::isdigit((unsigned char)a[1]))
            usage_error("Found argument '" + a + "' which wasn't expected, or isn't valid in this context");
        else if (!sub) { if (!m.values.count("input")) m.values["input"] = a; }   // value_of(): first value only
        else usage_error("Found argument '" + a + "' which wasn't expected, or isn't valid in this context");
    }
    return argc;
}

bool parse_f32(const std::string& s, float& out)
{
    char* end = nullptr;
    out = std::strtof(s.c_str(), &end);
    return !s.empty() && end && *end == '\0';
}

// match_turntable, inputs.rs:131-149
bool match_turntable(const Matches& m, float t[4], std::string& err)
{
    t[0] = t[1] = t[2] = 0.0f;
    const char* keys[3] = {"x", "y", "z"};
    for (int i = 0; i < 3; ++i) {
        auto it = m.values.find(keys[i]);
        if (it != m.values.end() && !parse_f32(it->second, t[i])) { err = "invalid float literal"; return false; }
    }
    t[3] = 1.0f;                               // no speed flag exists: 1.0 rad/s
    // test hook (tests/test_gpu_interactive.py): a standing turntable makes the interactive frames reproducible
    if (const char* s = std::getenv("SLOTH_SPEED")) t[3] = std::strtof(s, nullptr);
    t[1] += 3.14159265358979323846f;           // "All models for some reason are backwards"
    return true;
}

void check(int rc)
{
    if (rc != 0) {
        std::fprintf(stderr, "Error: %s\n", sloth_last_error());
        std::exit(1);
    }
}

termios g_saved;
volatile sig_atomic_t g_raw = 0;
void leave_raw()
{
    if (g_raw) {
        std::fputs("\x1b[?25h", stdout);       // cursor::Show
        std::fflush(stdout);
        tcsetattr(STDIN_FILENO, TCSANOW, &g_saved);
        g_raw = 0;
    }
}

// SIGTERM / SIGHUP / SIGINT (raw mode switches ISIG off, so these only arrive from outside): put the terminal back
// before dying -- only async-signal-safe calls here.
void on_fatal_signal(int sig)
{
    if (g_raw) {
        const char show[] = "\x1b[?25h";
        ssize_t ignored = write(STDOUT_FILENO, show, sizeof show - 1);
        (void)ignored;
        tcsetattr(STDIN_FILENO, TCSANOW, &g_saved);
        g_raw = 0;
    }
    _exit(128 + sig);
}

bool terminal_size(uint32_t& w, uint32_t& h)   // crossterm::terminal::size()
{
    winsize ws{};
    if (ioctl(STDOUT_FILENO, TIOCGWINSZ, &ws) != 0 || ws.ws_col == 0) return false;
    w = ws.ws_col;
    h = ws.ws_row;
    return true;
}

}  // namespace

int main(int argc, char** argv)
{
    Matches top, sub;
    int at = parse_level(argc, argv, 1, false, top);
    if (at < argc) parse_level(argc, argv, at + 1, true, sub);
    if (!top.values.count("input"))
        usage_error("The following required arguments were not provided:\n    <input filename(s)>...");
    if (sub.present && !sub.values.count("width"))
        usage_error("The following required arguments were not provided:\n    -w <width>");

    // match_meshes (main.rs:31) on the device: the files are parsed by sloth_scene_load and the soup never
    // visits the host.  Input that is legal but outside the device parser's grammar is parsed here instead.
    sloth_ctx* ctx = nullptr;
    check(sloth_ctx_create(0, sub.present ? 1 : 0, &ctx));
    std::string err;
    {
        const int rc = sloth_scene_load(ctx, top.values["input"].c_str(), nullptr, nullptr);
        if (rc == SLOTH_E_UNSUPPORTED) {
            std::fprintf(stderr, "note: %s -- parsing on the host\n", sloth_last_error());
            std::vector<sloth::SimpleMesh> meshes;
            if (!sloth::match_meshes(top.values["input"], meshes, err)) {
                std::fprintf(stderr, "Error: \"%s\"\n", err.c_str());
                return 1;
            }
            std::vector<float> xyz;
            std::vector<uint8_t> rgb;
            sloth::flatten(meshes, xyz, rgb);
            check(sloth_scene_set(ctx, xyz.data(), rgb.data(), rgb.size() / 3, sloth::scene_scale0(meshes)));
        } else if (rc != SLOTH_OK) {
            std::fprintf(stderr, "Error: \"%s\"\n", sloth_last_error());
            return 1;
        }
    }
    float turntable[4];
    if (!match_turntable(top, turntable, err)) { std::fprintf(stderr, "Error: %s\n", err.c_str()); return 1; }
    const bool no_color = top.no_color;        // main.rs:34 (top level only)
    const bool image = sub.present;
    bool webify = false;
    long webify_todo = 0;
    uint32_t W = 0, H = 0;
    if (image) {
        // match_dimensions, inputs.rs:159-169
        char* end = nullptr;
        W = (uint32_t)std::strtoul(sub.values["width"].c_str(), &end, 10);
        if (!end || *end) { std::fprintf(stderr, "Error: invalid digit found in string\n"); return 1; }
        H = W;
        if (sub.values.count("height")) {
            H = (uint32_t)std::strtoul(sub.values["height"].c_str(), &end, 10);
            if (!end || *end) { std::fprintf(stderr, "Error: invalid digit found in string\n"); return 1; }
        }
        if (!match_turntable(sub, turntable, err)) { std::fprintf(stderr, "Error: %s\n", err.c_str()); return 1; }
        if (sub.values.count("frame count")) {
            webify_todo = std::strtol(sub.values["frame count"].c_str(), &end, 10);   // i32 in the reference (main.rs:38-40)
            if (!end || *end) { std::fprintf(stderr, "Error: invalid digit found in string\n"); return 1; }
            if (webify_todo < 0) {
                // main.rs:57,92-99 with a negative count: the step is negative, the pitch never passes 9.42477 and
                // `count - 1 == frame` never holds -- the reference prints frames forever.  Refuse instead.
                std::fprintf(stderr, "Error: frame count %ld is negative (the reference would never stop exporting)\n", webify_todo);
                return 1;
            }
            webify = true;
        }
    }

    if (image) {
        check(sloth_ctx_resize(ctx, W, H));
        if (webify) {
            // main.rs:55-58,85-106: all frames are known up front -> one batched call
            const size_t cap = (size_t)(webify_todo > 0 ? webify_todo : 1) + 4;
            std::vector<float> pitches(cap);
            float y_arg = 0.0f;                // the helper adds PI itself (inputs.rs:148)
            auto it = sub.values.find("y");
            if (it != sub.values.end()) parse_f32(it->second, y_arg);
            const size_t n = sloth_turntable_pitches(y_arg, (uint32_t)webify_todo, pitches.data(), cap);
            std::vector<float> rots(n * 16);
            for (size_t k = 0; k < n; ++k) sloth_rotation_from_euler(turntable[0], pitches[k], turntable[2], &rots[k * 16]);
            std::fputs("let frames = [\n", stdout);
            // Context::flush runs on the GPU (sloth_render_text_batch): the host only frames the text
            const int mode = no_color ? 0 : 2;
            const size_t text_cap = sloth_text_capacity(ctx, mode);
            const size_t stride = (text_cap + 63) & ~(size_t)63;
            const size_t chunk = std::max<size_t>(1, std::min<size_t>(8, ((size_t)512 << 20) / stride));
            void* pinned = nullptr;
            check(sloth_pinned_alloc(chunk * stride, &pinned));
            std::vector<size_t> lens(chunk);
            for (size_t k0 = 0; k0 < n; k0 += chunk) {
                const size_t m = std::min(chunk, n - k0);
                check(sloth_render_text_batch(ctx, &rots[k0 * 16], m, mode, (char*)pinned, stride, lens.data()));
                for (size_t k = 0; k < m; ++k) {
                    std::fputs("`\n", stdout);
                    std::fwrite((char*)pinned + k * stride, 1, lens[k], stdout);
                    if (no_color) std::fputc('\n', stdout);            // println! of the plain frame
                    std::fputs((k0 + k == n - 1) ? "`];\n" : "`,\n", stdout);
                }
            }
            sloth_pinned_free(pinned);
        } else {
            float rot[16];
            sloth_rotation_from_euler(turntable[0], turntable[1], turntable[2], rot);
            const int mode = no_color ? 0 : 1;
            std::vector<char> text(sloth_text_capacity(ctx, mode) + 1);
            size_t len = 0;
            check(sloth_render_text(ctx, rot, mode, text.data(), text.size(), &len));
            std::fwrite(text.data(), 1, len, stdout);
            if (no_color) std::fputc('\n', stdout);
        }
        std::fflush(stdout);
        sloth_ctx_destroy(ctx);
        return 0;
    }

    // ---- interactive mode, main.rs:49-52,60-111 ------------------------------------------------
    if (tcgetattr(STDIN_FILENO, &g_saved) == 0) {
        termios raw = g_saved;
        cfmakeraw(&raw);
        tcsetattr(STDIN_FILENO, TCSANOW, &raw);
        g_raw = 1;
        std::atexit(leave_raw);
        struct sigaction sa;
        std::memset(&sa, 0, sizeof sa);
        sa.sa_handler = on_fatal_signal;
        sigaction(SIGTERM, &sa, nullptr);
        sigaction(SIGHUP, &sa, nullptr);
        sigaction(SIGINT, &sa, nullptr);
    }
    std::fputs("\x1b[?25l", stdout);           // cursor::Hide
    const double target_frame_time = 1.0 / 500.0;   // fps_cap = 500
    uint32_t cw = 0, ch = 0;
    std::vector<char> text;
    for (;;) {
        auto last = std::chrono::steady_clock::now();
        pollfd pfd{STDIN_FILENO, POLLIN, 0};
        if (poll(&pfd, 1, (int)(target_frame_time * 1000.0)) > 0) {
            char key = 0;
            const ssize_t got = read(STDIN_FILENO, &key, 1);
            if (got == 1 && (key == 'q' || key == 3)) break;   // 'q' or Ctrl-C
            if (got <= 0) break;   // end of input / hang-up: nobody can press q any more (and poll would spin)
        }
        uint32_t tw = 0, th = 0;
        if (!terminal_size(tw, th)) { leave_raw(); std::fprintf(stderr, "Error: cannot get the terminal size\n"); return 1; }
        if (tw != cw || th != ch) {            // Context::update adopts the terminal size, context.rs:134-137
            check(sloth_ctx_resize(ctx, tw, th));
            cw = tw;
            ch = th;
        }
        float rot[16];
        sloth_rotation_from_euler(turntable[0], turntable[1], turntable[2], rot);
        const int mode = no_color ? 0 : 1;
        text.resize(sloth_text_capacity(ctx, mode) + 1);
        size_t len = 0;
        check(sloth_render_text(ctx, rot, mode, text.data(), text.size(), &len));
        std::fputs("\x1b[1;1H", stdout);         // cursor::MoveTo(0,0), context.rs:53-55
        std::fwrite(text.data(), 1, len, stdout);
        if (no_color) std::fputc('\n', stdout);
        std::fflush(stdout);
        const float dt = (float)std::chrono::duration_cast<std::chrono::nanoseconds>(
                             std::chrono::steady_clock::now() - last).count() / 1000000000.0f;
        turntable[1] += turntable[3] * dt;     // main.rs:92-96
    }
    leave_raw();
    sloth_ctx_destroy(ctx);
    return 0;
}
