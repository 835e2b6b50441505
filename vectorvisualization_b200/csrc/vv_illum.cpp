// vv_illum.cpp -- host generation of the Zoeckler / Mallo illumination look-up tables
// (Illumination::createIllumTexZoeckler / createIllumTexMallo, VV/illumination.cpp:96-390; material = LineMat defaults,
// VV/illumination.h:40-62).  One-off, 256 x 256, so it stays on the host exactly as in the reference.
//
// init() calls createIllumTextures(true, true, true) but the float flag is not forwarded to the two generators
// (VV/illumination.cpp:59-91), so the textures are GL_LUMINANCE_ALPHA / GL_RGBA: the float tables are converted to
// UNORM8 on upload.  The tables returned here hold the decoded 8-bit values (b / 255).
#include <cmath>
#include <vector>

namespace vvb200 {

namespace {
const double kPi = 3.14159265358979323846;
// LineMat(), VV/illumination.h:42-54
const double kLight = 1.0, kAmbient = 0.1, kDiffuse = 0.5, kSpecular = 0.8, kDiffExp = 2.0;

float unorm8(double v)
{
    if (v < 0.0) v = 0.0;
    if (v > 1.0) v = 1.0;
    return (float)std::floor((float)v * 255.0f + 0.5f) / 255.0f;
}

// computeSpecTermIntegrandMallo, VV/illumination.cpp:381-390
double mallo_integrand(double beta, double n, double theta)
{
    double y = std::cos(theta - beta);
    if (y < 0.0) y = 0.0;
    return std::pow(y, n) * (std::cos(theta) / 2.0);
}

// computeSpecTermMallo, VV/illumination.cpp:352-378: composite Simpson rule, m = 10
double mallo_spec_term(double alpha, double beta, double n)
{
    const double a = alpha - kPi / 2.0, b = kPi / 2.0;
    const int m = 10;
    const double h = (b - a) / (2.0 * m);
    double integral = 0.0;
    for (int i = 0; i < 2 * m; i += 2) {
        integral += 2.0 * mallo_integrand(beta, n, a + i * h);
        integral += 4.0 * mallo_integrand(beta, n, a + (i + 1) * h);
    }
    integral += mallo_integrand(beta, n, b);
    integral -= mallo_integrand(beta, n, a);   // f(a) was counted twice inside the loop
    return integral * (h / 3.0);
}
} // namespace

void make_illum_tables(int w, int h, float spec_exp, std::vector<float> &zoeckler, std::vector<float> &mallo_diff,
                       std::vector<float> &mallo_spec)
{
    zoeckler.resize((size_t)2 * w * h);
    mallo_diff.resize((size_t)w * h);
    mallo_spec.resize((size_t)w * h);
    const double irx = 1.0 / (double)(w - 1), iry = 1.0 / (double)(h - 1);
    size_t zi = 0, mi = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            {   // Zoeckler et al. VIS 96, VV/illumination.cpp:121-171: texel grid spans [0,1] inclusively
                const double lt = 2.0 * ((double)x * irx) - 1.0, vt = 2.0 * ((double)y * iry) - 1.0;
                double diffuse = std::sqrt(1.0 - lt * lt);
                diffuse = std::pow(diffuse, kDiffExp);
                double dp = lt * vt - std::sqrt(1.0 - lt * lt) * std::sqrt(1.0 - vt * vt);
                dp = (dp < -1.0) ? -1.0 : ((dp > 1.0) ? 1.0 : dp);
                double td = kAmbient * kLight + diffuse * kDiffuse * kLight;
                double ts = std::pow(std::fabs(dp), (double)spec_exp) * kSpecular * kLight;
                td *= 0.5; td = (td < 0.0) ? 0.0 : ((td > 1.0) ? 1.0 : td);
                ts *= 0.5; ts = (ts < 0.0) ? 0.0 : ((ts > 1.0) ? 1.0 : ts);
                zoeckler[zi++] = unorm8((double)(float)td);
                zoeckler[zi++] = unorm8((double)((float)ts * 0.9f));
            }
            {   // Mallo et al. VIS 2005, VV/illumination.cpp:238-271: texel centres
                const double s = ((double)x + 0.5) / w, t = ((double)y + 0.5) / h;
                const double alpha = std::acos(2.0 * s - 1.0), beta = std::acos(2.0 * t - 1.0);
                const double lt = 2.0 * t - 1.0;
                const double diffuse = std::sqrt(1.0 - lt * lt) * (std::sin(alpha) + (kPi - alpha) * std::cos(alpha)) * 0.25;
                const double specular = 3.5 * mallo_spec_term(alpha, beta, (double)spec_exp);
                double c = diffuse * kDiffuse * kLight;
                c = (c < 0.0) ? 0.0 : ((c > 1.0) ? 1.0 : c);
                mallo_diff[mi] = unorm8((double)(float)c);
                c = specular * kSpecular * kLight;
                c = (c < 0.0) ? 0.0 : ((c > 1.0) ? 1.0 : c);
                mallo_spec[mi] = unorm8((double)(float)c);
                ++mi;
            }
        }
}

} // namespace vvb200

#include "../../include/vv_c_api.h"

extern "C" int vv_make_illum_tables(float spec_exp, int width, int height, float *zoeckler_la, float *mallo_diffuse, float *mallo_specular)
{
    if (width < 2 || height < 2 || !zoeckler_la || !mallo_diffuse || !mallo_specular) return VV_ERR_INVALID;
    std::vector<float> z, d, s;
    vvb200::make_illum_tables(width, height, spec_exp, z, d, s);
    for (size_t i = 0; i < z.size(); ++i) zoeckler_la[i] = z[i];
    for (size_t i = 0; i < d.size(); ++i) { mallo_diffuse[i] = d[i]; mallo_specular[i] = s[i]; }
    return VV_OK;
}
