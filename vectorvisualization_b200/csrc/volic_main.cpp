// volic -- headless clone of the reference's executable (VV/3DLIC.cpp main/init, VV/README.txt:11-75):
//
//   volic <volfilename.dat> [-g] [-f <png>] [-n <noisefile>] [-t <png>]            (the reference's own argv grammar)
//         [--out=<png>] [--size=WxH] [--scalar=<dat>] [--illum=gradient|mallo|zoeckler|none] [--technique=raycast|licvolume]
//         [--step=<stepSizeVol>] [--lic-step=<stepSizeLIC>] [--steps=<fwd>,<bwd>] [--freq=<f>] [--tf-mode=b|a|r|length|scalar]
//
// GLUT / GLEW are replaced by a single frame written as PNG (default snapshotOut_snapshot.png, the name pattern of
// VV/renderer.cpp:1487).  Extra options mirror the keyboard bindings of VV/3DLIC.cpp:243-488.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vv_c_api.h"

#define CHECK(x)                                                                       \
    do {                                                                               \
        if ((x) != VV_OK) {                                                            \
            std::fprintf(stderr, "volic: %s\n", vv_last_error());                     \
            return 1;                                                                  \
        }                                                                              \
    } while (0)

int main(int argc, char **argv)
{
    std::vector<const char *> ref_args;
    std::string out = "snapshotOut_snapshot.png", scalar, illum = "none", technique = "raycast", tfmode = "b";
    int w = 1280, h = 960;                                  // WINDOW_WIDTH x WINDOW_HEIGHT, VV/types.h:45-46
    VVLicParams lp;
    vv_default_lic_params(&lp);
    ref_args.push_back(argv[0]);
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (!std::strncmp(a, "--out=", 6)) out = a + 6;
        else if (!std::strncmp(a, "--size=", 7)) { if (std::sscanf(a + 7, "%dx%d", &w, &h) != 2) { std::fprintf(stderr, "bad --size\n"); return 1; } }
        else if (!std::strncmp(a, "--scalar=", 9)) scalar = a + 9;
        else if (!std::strncmp(a, "--illum=", 8)) illum = a + 8;
        else if (!std::strncmp(a, "--technique=", 12)) technique = a + 12;
        else if (!std::strncmp(a, "--step=", 7)) lp.stepSizeVol = (float)std::atof(a + 7);
        else if (!std::strncmp(a, "--lic-step=", 11)) lp.stepSizeLIC = (float)std::atof(a + 11);
        else if (!std::strncmp(a, "--steps=", 8)) { if (std::sscanf(a + 8, "%d,%d", &lp.stepsForward, &lp.stepsBackward) != 2) { std::fprintf(stderr, "bad --steps\n"); return 1; } }
        else if (!std::strncmp(a, "--freq=", 7)) lp.freqScale = (float)std::atof(a + 7);
        else if (!std::strncmp(a, "--tf-mode=", 10)) tfmode = a + 10;
        else ref_args.push_back(a);
    }
    VVArgs args;
    if (vv_parse_args((int)ref_args.size(), ref_args.data(), &args) != VV_OK || args.show_help || !args.vol_file[0]) {
        if (!args.show_help) std::fprintf(stderr, "%s\n", vv_last_error());
        std::fputs(vv_usage(), stderr);                       // VV/3DLIC.cpp:848-852
        return args.show_help ? 0 : 1;
    }
    VVRenderer *r = nullptr;
    CHECK(vv_create(&r, 0));
    std::string defines;
    if (illum == "gradient") defines = "#define ILLUM_GRADIENT";
    else if (illum == "mallo") defines = "#define ILLUM_MALLO";
    else if (illum == "zoeckler") defines = "#define ILLUM_ZOECKLER";
    CHECK(vv_init(r, defines.empty() ? nullptr : defines.c_str()));
    CHECK(vv_load_dat(r, args.vol_file));                                           // VV/3DLIC.cpp:686-707
    if (args.noise_file[0]) CHECK(vv_load_noise(r, args.noise_file, args.use_gradients));
    else CHECK(vv_generate_white_noise(r, 256, 0, 1.0f / 6.0f, args.use_gradients));   // VV/dataset.cpp:1142-1163
    if (!scalar.empty()) CHECK(vv_load_scalar_dat(r, scalar.c_str()));              // hard-coded path in VV/3DLIC.cpp:717
    else CHECK(vv_set_option(r, VV_OPT_NOISE_GATE, 0));
    if (args.filter_file[0]) {
        if (vv_load_filter_png(r, args.filter_file) != VV_OK) {                       // VV/3DLIC.cpp:725-731
            std::fprintf(stderr, "could not load lic filter kernel ... using box filter\n");
            CHECK(vv_set_box_filter(r, 256));
        }
    }
    if (args.tf_file[0] && vv_load_tf_png(r, args.tf_file) != VV_OK) std::fprintf(stderr, "could not load transfer function\n");
    const char *modes[] = {"b", "a", "r", "length", "scalar"};
    for (int i = 0; i < 5; ++i) if (tfmode == modes[i]) CHECK(vv_set_option(r, VV_OPT_TF_MODE, i));
    CHECK(vv_set_lic_params(r, &lp));
    CHECK(vv_set_technique(r, technique == "licvolume" ? VV_VOLIC_LICVOLUME : VV_VOLIC_RAYCAST));
    CHECK(vv_update_light_pos(r));
    CHECK(vv_resize(r, w, h));
    CHECK(vv_render(r, 1));
    CHECK(vv_save_png(r, out.c_str(), 1));
    std::printf("volic: %dx%d, %llu ray samples, dominant kernel %.3f ms -> %s\n", w, h, (unsigned long long)vv_last_ray_samples(r),
                vv_last_kernel_ms(r), out.c_str());
    vv_destroy(r);
    return 0;
}
