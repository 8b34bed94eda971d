// Per-handle tuning options of libfdtd2d (host only).  A handle copies the defaults when it is created -- the defaults
// come from the FDTD2D_* environment variables, read ONCE at that moment -- and fdtd2d_set_option changes them for that
// handle alone afterwards, so nothing on the stepping path calls getenv() and two host threads driving two handles
// never share mutable state.
#pragma once
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>

namespace fdtd2d {

struct Options {
    int wavefront = 1;         // k = 8 / 12 passes (fp64: 4 / 6 / 8) of large grids on the row-streaming wavefront kernel
    int wave_min_tiles = -1;   // plain fp32-sized tiles (48 x 112 cells) from which the wavefront takes over; -1: 0.4 per warp of the GPU
    int ring_min_tiles = -1;   // ... from which the left / right Mur ring rides along the wavefront; -1: 2 per warp
    int ring_strips = 1;       // left / right ring strips on the wavefront at all
    int wave_run_rows = 640;   // cap on the rows of one wavefront run
    int edge_reserve = -1;     // whole grids: SMs the wavefront kernel leaves to the edge tiles of its pass (the runs are cut for the
                               // other SMs and launched first); 0 none (edge tiles first, wavefront on every SM), -1: automatic --
                               // when the edge tiles are 2 .. 25 % of the pass (4096^2: 21 SMs, 16384^2: 6)
    int ring_cost = 0;         // cost of a ring-strip row in percent of a plain row when runs are balanced, warm-up rows included;
                               // 0: 208 / 220 with reserved SMs, else the older rule (ring runs half as long as plain runs)
    int auto_k12 = 0;          // k_temporal = 0 picks the 12-level wavefront when it exists (uniform permeability)
    int uniform_ch = 1;        // pass dt/(mu*dx) as a scalar when the map is uniform
    int resident = 1;          // cluster-resident kernel for small fp32 grids
    int resident_cfg = 5;      // its shape (index into kResCfgs): 5 = the packed kernel (grid_resident_x2.cuh; grids it does not take fall to 0)
    int resident_cluster = 0;  // CTAs per grid (0: as few as fit)
    int resident_trim = -1;    // rows the first / last CTA of a cluster hold fewer than the others (-1: 4 x rows per thread);
                               // packed kernel: rows of the last band (<= 0: automatic)
    int resident_rows = 0;     // packed kernel: rows of the middle bands (0: automatic)
    int tma_pair = 0;          // pairwise mbarriers instead of CTA barriers in the TMA tile kernel
    int f64_k = 0;             // default k_temporal of fp64 handles (0: 8 on the wavefront, else 4)
    int fuse = 0;              // EXPERIMENT, off: two k = 8 passes per launch, the second reading the first's output from L2
                               // (1 on, -1 on for large grids).  Bit-exact but slower on B200, see DESIGN.md 9.
    int stage = 0;             // k = 12 passes on the staged wavefront (three warps of 4 levels each, strip_stage.cuh) instead of
                               // the one-warp 12-level instance; 4 / 5 = groups per CTA
    int debug = 0;             // print launch geometry to stderr
    int measure_skip = 0;      // MEASUREMENT AID, results are wrong: 1 = passes launch no edge tiles, 2 = no runs / plain tiles,
                               // 4 = edge tiles on the main stream (no overlap with the runs)
};

struct OptionKey {
    const char* name;
    int Options::*field;
};

inline const OptionKey* option_keys(int* n) {
    static const OptionKey keys[] = {
        {"wavefront", &Options::wavefront},
        {"wave_min_tiles", &Options::wave_min_tiles},
        {"ring_min_tiles", &Options::ring_min_tiles},
        {"ring_strips", &Options::ring_strips},
        {"wave_run_rows", &Options::wave_run_rows},
        {"edge_reserve", &Options::edge_reserve},
        {"ring_cost", &Options::ring_cost},
        {"auto_k12", &Options::auto_k12},
        {"uniform_ch", &Options::uniform_ch},
        {"resident", &Options::resident},
        {"resident_cfg", &Options::resident_cfg},
        {"resident_cluster", &Options::resident_cluster},
        {"resident_trim", &Options::resident_trim},
        {"resident_rows", &Options::resident_rows},
        {"tma_pair", &Options::tma_pair},
        {"f64_k", &Options::f64_k},
        {"fuse", &Options::fuse},
        {"stage", &Options::stage},
        {"debug", &Options::debug},
        {"measure_skip", &Options::measure_skip},
    };
    *n = (int)(sizeof(keys) / sizeof(keys[0]));
    return keys;
}

inline const OptionKey* find_option(const char* name) {
    int n = 0;
    const OptionKey* keys = option_keys(&n);
    for (int i = 0; i < n; ++i)
        if (name && strcmp(name, keys[i].name) == 0) return &keys[i];
    return nullptr;
}

// Defaults for a new handle: FDTD2D_<KEY> for every key above, plus the negative spellings of round 1
// (FDTD2D_NO_RESIDENT, FDTD2D_NO_RING_STRIPS, FDTD2D_NO_UNIFORM_CH).
inline Options options_from_env() {
    Options o;
    int n = 0;
    const OptionKey* keys = option_keys(&n);
    for (int i = 0; i < n; ++i) {
        std::string env = "FDTD2D_";
        for (const char* c = keys[i].name; *c; ++c) env += (char)toupper((unsigned char)*c);
        if (const char* v = getenv(env.c_str()))
            if (*v) o.*(keys[i].field) = atoi(v);
    }
    auto on = [](const char* name) {
        const char* v = getenv(name);
        return v && atoi(v) != 0;
    };
    if (on("FDTD2D_NO_RESIDENT")) o.resident = 0;
    if (on("FDTD2D_NO_RING_STRIPS")) o.ring_strips = 0;
    if (on("FDTD2D_NO_UNIFORM_CH")) o.uniform_ch = 0;
    return o;
}

}  // namespace fdtd2d
