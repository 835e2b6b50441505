// vv_io.cpp -- headless replacements of the reference's file-facing surface:
//   DatFile           VV/reader.cpp:81-336        (.dat / .raw)
//   NoiseDataSet      VV/dataset.cpp:1347-1389    (3 x int32 + u8 noise file)
//   pngRead/pngWrite  VV/imageUtils.cpp:19-347    (8-bit gray / GA / RGB / RGBA) -- zlib only, no libpng
//   TransferEdit      VV/transferEdit.cpp:113-337 (<name>_rgba.png + <name>_alpha.png)
//   ParseArguments    VV/parseArg.cpp:97-365
#include "vv_host.h"

#include <zlib.h>

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vvb200 {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg) { g_error = msg; return code; }
const std::string &last_error_string() { return g_error; }

int next_pow2(int v) { int i = 1; while (i < v) i <<= 1; return i; }   // VV/mmath.cpp:37-42

// ------------------------------------------------------------------------------------------------ PNG
static uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static void put32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

static int paeth(int a, int b, int c)
{
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return (pb <= pc) ? b : c;
}

static bool png_read_impl(const char *path, std::vector<uint8_t> &out, int &w, int &h, int &channels, std::string &err);

// no exception leaves this function: its callers are extern "C" entry points (a hostile header could otherwise size a vector
// the allocator refuses)
bool png_read_file(const char *path, std::vector<uint8_t> &out, int &w, int &h, int &channels, std::string &err)
{
    try {
        return png_read_impl(path, out, w, h, channels, err);
    } catch (const std::exception &e) {
        err = std::string("PNG read failed: ") + e.what();
        return false;
    }
}

static bool png_read_impl(const char *path, std::vector<uint8_t> &out, int &w, int &h, int &channels, std::string &err)
{
    FILE *fp = std::fopen(path, "rb");
    if (!fp) { err = std::string("Could not open PNG file ") + path; return false; }
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof(tmp), fp)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(fp);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) { err = "not a PNG file"; return false; }
    size_t pos = 8;
    int bitDepth = 0, colorType = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    w = h = 0;
    while (pos + 12 <= buf.size()) {
        uint32_t len = be32(&buf[pos]);
        const char *type = (const char *)&buf[pos + 4];
        if (pos + 12 + len > buf.size()) { err = "truncated PNG chunk"; return false; }
        const uint8_t *d = &buf[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad PNG header"; return false; }
            const uint32_t uw = be32(d), uh = be32(d + 4);
            if (uw == 0 || uh == 0 || uw > 65536u || uh > 65536u) { err = "PNG dimensions out of range"; return false; }
            w = (int)uw; h = (int)uh; bitDepth = d[8]; colorType = d[9]; interlace = d[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (w <= 0 || h <= 0) { err = "PNG without header"; return false; }
    if (idat.empty()) { err = "PNG without image data"; return false; }
    if (bitDepth != 8) { err = "pngRead: currently only 8 Bit per channel are supported."; return false; }   // imageUtils.cpp:71-77
    if (interlace) { err = "pngRead: interlaced PNG not supported"; return false; }
    int srcCh;
    switch (colorType) {
    case 0: srcCh = 1; break;
    case 4: srcCh = 2; break;
    case 2: srcCh = 3; break;
    case 6: srcCh = 4; break;
    case 3: srcCh = 1; break;   // palette: expanded to RGB below
    default: err = "pngRead: invalid color type"; return false;
    }
    const size_t stride = (size_t)w * srcCh;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawLen = raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), idat.size()) != Z_OK || rawLen != raw.size()) { err = "PNG inflate failed"; return false; }
    std::vector<uint8_t> img(stride * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *src = &raw[(stride + 1) * y];
        uint8_t *dst = &img[stride * y];
        const uint8_t *up = y ? &img[stride * (y - 1)] : nullptr;
        const int ft = src[0];
        for (size_t i = 0; i < stride; ++i) {
            int a = (i >= (size_t)srcCh) ? dst[i - srcCh] : 0;
            int b = up ? up[i] : 0;
            int c = (up && i >= (size_t)srcCh) ? up[i - srcCh] : 0;
            int x = src[1 + i];
            switch (ft) {
            case 0: break;
            case 1: x += a; break;
            case 2: x += b; break;
            case 3: x += (a + b) >> 1; break;
            case 4: x += paeth(a, b, c); break;
            default: err = "bad PNG filter"; return false;
            }
            dst[i] = (uint8_t)x;
        }
    }
    if (colorType == 3) {
        channels = 3;
        out.resize((size_t)w * h * 3);
        for (size_t i = 0; i < (size_t)w * h; ++i)
            for (int k = 0; k < 3; ++k) out[3 * i + k] = (3 * (size_t)img[i] + k < plte.size()) ? plte[3 * img[i] + k] : 0;
    } else {
        channels = srcCh;
        out.swap(img);
    }
    return true;
}

static void png_chunk(std::vector<uint8_t> &f, const char *type, const std::vector<uint8_t> &data)
{
    put32(f, (uint32_t)data.size());
    size_t s = f.size();
    f.insert(f.end(), type, type + 4);
    f.insert(f.end(), data.begin(), data.end());
    put32(f, (uint32_t)crc32(0L, &f[s], (uInt)(f.size() - s)));
}

bool png_write_file(const char *path, const uint8_t *data, int w, int h, int channels, std::string &err)
{
    if (channels < 1 || channels > 4) { err = "pngWrite: 1..4 channels"; return false; }
    static const int ctype[5] = {0, 0, 4, 2, 6};
    std::vector<uint8_t> f = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put32(ihdr, w); put32(ihdr, h);
    ihdr.push_back(8); ihdr.push_back(ctype[channels]); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(f, "IHDR", ihdr);
    const size_t stride = (size_t)w * channels;
    std::vector<uint8_t> raw((stride + 1) * h);
    for (int y = 0; y < h; ++y) {
        raw[(stride + 1) * y] = 0;
        std::memcpy(&raw[(stride + 1) * y + 1], data + stride * y, stride);
    }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) { err = "PNG deflate failed"; return false; }
    comp.resize(clen);
    png_chunk(f, "IDAT", comp);
    png_chunk(f, "IEND", {});
    FILE *fp = std::fopen(path, "wb");
    if (!fp) { err = std::string("cannot write ") + path; return false; }
    std::fwrite(f.data(), 1, f.size(), fp);
    std::fclose(fp);
    return true;
}

// ------------------------------------------------------------------------------------------------ DAT
static int type_size(int t) { return t == VV_UCHAR ? 1 : (t == VV_USHORT ? 2 : (t == VV_FLOAT ? 4 : 0)); }

size_t dat_bytes(const VVDatInfo *i)
{
    return (size_t)type_size(i->data_type) * i->data_dim * i->resolution[0] * i->resolution[1] * i->resolution[2];
}

static bool file_exists(const char *p)
{
    FILE *f = std::fopen(p, "rb");
    if (!f) return false;
    std::fclose(f);
    return true;
}

// DatFile::parseDatFile, VV/reader.cpp:81-263.  Same keys, same '#' comment rule, same raw-file search
// (as given, else relative to the .dat's directory), same TimeDependent range check.
int parse_dat(const char *path, VVDatInfo *out)
{
    if (!path || !out) return fail(VV_ERR_INVALID, "DatFile: null argument");
    std::memset(out, 0, sizeof(*out));
    out->slice_thickness[0] = out->slice_thickness[1] = out->slice_thickness[2] = 1.0f;
    out->data_dim = 1;
    FILE *fp = std::fopen(path, "rb");
    if (!fp) return fail(VV_ERR_IO, std::string("opening .dat file failed: ") + path);
    char line[255], raw[255] = "";
    bool parseError = false, haveRaw = false;
    while (std::fgets(line, sizeof(line), fp)) {
        if (line[0] == '#') continue;
        char *cp;
        if (std::strstr(line, "ObjectFileName")) {
            if (!(cp = std::strchr(line, ':')) || std::sscanf(cp + 1, "%254s", raw) != 1) { parseError = true; break; }
            haveRaw = true;
        } else if (std::strstr(line, "Resolution")) {
            if (!(cp = std::strchr(line, ':')) ||
                std::sscanf(cp + 1, "%i %i %i", &out->resolution[0], &out->resolution[1], &out->resolution[2]) != 3) { parseError = true; break; }
        } else if (std::strstr(line, "SliceThickness")) {
            if (!(cp = std::strchr(line, ':')) ||
                std::sscanf(cp + 1, "%f %f %f", &out->slice_thickness[0], &out->slice_thickness[1], &out->slice_thickness[2]) != 3) { parseError = true; break; }
        } else if (std::strstr(line, "Format")) {
            if ((cp = std::strstr(line, "UCHAR"))) out->data_type = VV_UCHAR;
            else if ((cp = std::strstr(line, "USHORT"))) out->data_type = VV_USHORT;
            else if ((cp = std::strstr(line, "FLOAT"))) out->data_type = VV_FLOAT;
            else { std::fclose(fp); return fail(VV_ERR_IO, "DatFile:  Cannot process data other than of UCHAR*, USHORT* and FLOAT* format."); }
            // parseDataDim, reader.cpp:308-325: first digit run after the type name, default 1
            while (*cp && !std::isdigit((unsigned char)*cp)) ++cp;
            if (!*cp || std::sscanf(cp, "%i", &out->data_dim) != 1) out->data_dim = 1;
        } else if (std::strstr(line, "TimeDependent")) {
            if (!(cp = std::strchr(line, ':')) || std::sscanf(cp + 1, "%i %i", &out->time_begin, &out->time_end) != 2) { parseError = true; break; }
            if (out->time_begin < 0 || out->time_end < out->time_begin) {
                std::fclose(fp);
                return fail(VV_ERR_IO, "DatFile:  Illegal boundaries for Timedependent Data Set");
            }
        }
        // other lines are skipped (reader.cpp:203-206)
    }
    std::fclose(fp);
    if (parseError) return fail(VV_ERR_IO, std::string("parse error: ") + line);
    if (!haveRaw) return fail(VV_ERR_IO, "DatFile:  no ObjectFileName");
    char name[512];
    std::snprintf(name, sizeof(name), raw, out->time_begin);
    std::string pattern = raw;
    if (!file_exists(name)) {
        std::string dat = path;
        size_t s = dat.find_last_of('/');
        if (s == std::string::npos) s = dat.find_last_of('\\');
        if (s == std::string::npos) return fail(VV_ERR_IO, std::string("DatFile:  No valid search path for RAW file \"") + raw + "\".");
        pattern = dat.substr(0, s + 1) + raw;
        std::snprintf(name, sizeof(name), pattern.c_str(), out->time_begin);
        if (!file_exists(name)) return fail(VV_ERR_IO, std::string("DatFile:  Could not open RAW file \"") + name + "\".");
    }
    for (int t = out->time_begin + 1; t <= out->time_end; ++t) {
        std::snprintf(name, sizeof(name), pattern.c_str(), t);
        if (!file_exists(name)) return fail(VV_ERR_IO, std::string("DatFile:  Could not open RAW file \"") + name + "\" for timestep.");
    }
    if (pattern.size() >= sizeof(out->raw_file)) return fail(VV_ERR_IO, "DatFile: raw path too long");
    std::strcpy(out->raw_file, pattern.c_str());
    return VV_OK;
}

// DatFile::readRawData, VV/reader.cpp:266-305
int read_raw(const VVDatInfo *info, int time_step, void *out, size_t out_bytes)
{
    if (!info || !out) return fail(VV_ERR_INVALID, "readRawData: null argument");
    if (time_step < info->time_begin || time_step > info->time_end) return fail(VV_ERR_INVALID, "readRawData: time step out of range");
    const size_t need = dat_bytes(info);
    if (out_bytes < need) return fail(VV_ERR_INVALID, "readRawData: output buffer too small");
    char name[600];
    std::snprintf(name, sizeof(name), info->raw_file, time_step);
    FILE *fp = std::fopen(name, "rb");
    if (!fp) return fail(VV_ERR_IO, std::string("Could not open RAW file. No file \"") + name + "\".");
    size_t got = std::fread(out, 1, need, fp);
    std::fclose(fp);
    if (got != need) return fail(VV_ERR_IO, std::string("Reading volume data \"") + name + "\" failed.");
    return VV_OK;
}

// NoiseDataSet::loadRawData, VV/dataset.cpp:1347-1389
int read_noise_file(const char *path, std::vector<uint8_t> &data, int dims[3])
{
    FILE *fp = path ? std::fopen(path, "rb") : nullptr;
    if (!fp) return fail(VV_ERR_IO, std::string("NoiseData:  Could not load noise from (\"") + (path ? path : "") + "\").");
    int32_t hdr[3];
    if (std::fread(hdr, 4, 3, fp) != 3) { std::fclose(fp); return fail(VV_ERR_IO, "NoiseData:  Could not read noise header"); }
    if (hdr[0] <= 0 || hdr[1] <= 0 || hdr[2] <= 0 || (int64_t)hdr[0] * hdr[1] * hdr[2] > ((int64_t)1 << 33)) {
        std::fclose(fp);
        return fail(VV_ERR_IO, "NoiseData:  bad noise header");
    }
    size_t n = (size_t)hdr[0] * hdr[1] * hdr[2];
    data.resize(n);
    size_t got = std::fread(data.data(), 1, n, fp);
    std::fclose(fp);
    if (got != n) return fail(VV_ERR_IO, "NoiseData:  Error reading noise data");
    dims[0] = hdr[0]; dims[1] = hdr[1]; dims[2] = hdr[2];
    return VV_OK;
}

// TransferEdit::setTFFileNames + loadTF, VV/transferEdit.cpp:113-160, 224-337
int load_tf_png(const char *name, uint8_t *tf)
{
    if (!name) return fail(VV_ERR_INVALID, "TransferEdit Load:  No filename set.");
    std::string base = name, ext;
    size_t dot = base.rfind('.');
    if (dot != std::string::npos) { ext = base.substr(dot); base = base.substr(0, dot); } else ext = ".png";
    const std::string fRGBA = base + "_rgba" + ext, fAO = base + "_alpha" + ext;
    std::vector<uint8_t> img;
    int w, h, ch;
    std::string err;
    bool okRGBA = false, okAO = false;
    if (png_read_file(fRGBA.c_str(), img, w, h, ch, err) && w * h >= 256) {
        okRGBA = true;
        for (int i = 0; i < 256; ++i) {
            if (ch < 3) {
                tf[5 * i] = tf[5 * i + 1] = tf[5 * i + 2] = img[i * ch];
                if (ch == 2) tf[5 * i + 3] = img[i * ch + 1];
            } else {
                tf[5 * i] = img[i * ch]; tf[5 * i + 1] = img[i * ch + 1]; tf[5 * i + 2] = img[i * ch + 2];
                if (ch == 4) tf[5 * i + 3] = img[i * ch + 3];
            }
        }
    }
    // NB: `!loadRGBATF() && !loadAlphaOpacTF()` short-circuits: the alpha file is only read when the RGBA file failed
    // (VV/transferEdit.cpp:239)
    if (!okRGBA) {
        if (png_read_file(fAO.c_str(), img, w, h, ch, err) && w * h >= 256) {
            okAO = true;
            for (int i = 0; i < 256; ++i) {
                tf[5 * i + 3] = img[i * ch];
                if (ch == 2) tf[5 * i + 4] = img[i * ch + 1];
            }
        }
    }
    if (!okRGBA && !okAO) return fail(VV_ERR_IO, "could not load transfer function: " + fRGBA + " / " + fAO);
    return VV_OK;
}

} // namespace vvb200

// ------------------------------------------------------------------------------------------------ C ABI (IO part)
using namespace vvb200;

extern "C" {

const char *vv_last_error(void)
{
    return vvb200::last_error_string().c_str();
}

// loadGradients / saveGradients with DATRAW_UCHAR, VV/gradient.cpp:93-187
int vv_grd_read(const char *file_name, const int dims[3], uint8_t *gradients3)
{
    if (!file_name || !dims || !gradients3) return fail(VV_ERR_INVALID, "vv_grd_read: null argument");
    const std::string p = std::string(file_name) + ".grd";
    FILE *fp = std::fopen(p.c_str(), "rb");
    if (!fp) return fail(VV_ERR_IO, "loadGradients: No pre-computed gradients found.");
    const size_t n = (size_t)3 * dims[0] * dims[1] * dims[2];
    const size_t got = std::fread(gradients3, 1, n, fp);
    std::fclose(fp);
    if (got != n) return fail(VV_ERR_IO, "loadGradients: Reading gradients from \"" + p + "\" failed.");
    return VV_OK;
}

int vv_grd_write(const char *file_name, const int dims[3], const uint8_t *gradients3)
{
    if (!file_name || !dims || !gradients3) return fail(VV_ERR_INVALID, "vv_grd_write: null argument");
    const std::string p = std::string(file_name) + ".grd";
    FILE *fp = std::fopen(p.c_str(), "wb");
    if (!fp) return fail(VV_ERR_IO, "saveGradients: Could not open file \"" + p + "\".");
    const size_t n = (size_t)3 * dims[0] * dims[1] * dims[2];
    const size_t put = std::fwrite(gradients3, 1, n, fp);
    const int rc = std::fclose(fp);
    if (put != n || rc != 0) return fail(VV_ERR_IO, "saveGradients: Writing gradients to \"" + p + "\" failed.");
    return VV_OK;
}

int vv_parse_dat(const char *dat_path, VVDatInfo *out) { return parse_dat(dat_path, out); }
int vv_read_raw(const VVDatInfo *info, int time_step, void *out, size_t out_bytes) { return read_raw(info, time_step, out, out_bytes); }

int vv_png_read(const char *path, uint8_t **data, int *w, int *h, int *channels)
{
    if (!path || !data || !w || !h || !channels) return fail(VV_ERR_INVALID, "vv_png_read: null argument");
    std::vector<uint8_t> img;
    std::string err;
    if (!png_read_file(path, img, *w, *h, *channels, err)) return fail(VV_ERR_IO, err);
    *data = (uint8_t *)std::malloc(img.size());
    if (!*data) return fail(VV_ERR_IO, "vv_png_read: out of memory");
    std::memcpy(*data, img.data(), img.size());
    return VV_OK;
}
void vv_free(void *p) { std::free(p); }
int vv_png_write(const char *path, const uint8_t *data, int w, int h, int channels)
{
    std::string err;
    if (!png_write_file(path, data, w, h, channels, err)) return fail(VV_ERR_IO, err);
    return VV_OK;
}

static const char *kUsage =
    "\nUsage:  volic <volfilename.dat> [-h | --help] [-g | --gradient] \n"
    "\t\t\t\t[-f <file> | --filter=<file>]\n"
    "\t\t\t\t[-n <file> | --noise=<file>]\n"
    "\t\t\t\t[-t <file> | --transfer=<file>]\n"
    "\t-h | --help \tShow usage\n"
    "\t-g | --gradient\tUse noise gradients\n"
    "\t-f <png>\tFilter kernel stored in PNG file\n"
    "\t--filter=<png>\n"
    "\t-n <noisefile>\tUse given noise for LIC\n"
    "\t--noise=<noisefile>\n"
    "\t-t <png>\tTransfer function stored in PNG file\n"
    "\t--transfer=<png>\n";
const char *vv_usage(void) { return kUsage; }

static bool take(char *dst, const char *src)
{
    if (std::strlen(src) >= 512) return false;
    std::strcpy(dst, src);
    return true;
}

// ParseArguments::parse / parseLongArgs, VV/parseArg.cpp:97-365
int vv_parse_args(int argc, const char *const *argv, VVArgs *out)
{
    if (!out || !argv) return fail(VV_ERR_INVALID, "vv_parse_args: null argument");
    std::memset(out, 0, sizeof(*out));
    if (argc == 1) return fail(VV_ERR_INVALID, "no arguments");                       // :102-105
    for (int idx = 1; idx < argc; ++idx) {
        const char *a = argv[idx];
        const int len = (int)std::strlen(a);
        if (len > 2 && a[0] == '-' && a[1] == '-') {                                  // long form, :110-117, 277-365
            const char *k = a + 2;
            bool ok = true;
            if (!std::strcmp(k, "help")) out->show_help = 1;
            else if (!std::strncmp(k, "filter", 6)) { if (len > 9 && a[8] == '=') ok = take(out->filter_file, a + 9); else return fail(VV_ERR_INVALID, "Missing filename:  filter kernel (png)"); }
            else if (!std::strncmp(k, "noise", 5)) { if (len > 9 && a[7] == '=') ok = take(out->noise_file, a + 8); else return fail(VV_ERR_INVALID, "Missing filename:  noise"); }
            else if (!std::strncmp(k, "transfer", 8)) { if (len > 11 && a[10] == '=') ok = take(out->tf_file, a + 11); else return fail(VV_ERR_INVALID, "Missing filename:  transfer function (png)"); }
            else if (!std::strncmp(k, "redirect", 8)) { if (len > 11 && a[10] == '=') ok = take(out->redirect_file, a + 11); else return fail(VV_ERR_INVALID, "Missing filename:  redirection"); }
            else if (!std::strncmp(k, "halton", 6)) { if (len > 9 && a[8] == '=') ok = take(out->halton_file, a + 9); else return fail(VV_ERR_INVALID, "Missing filename:  halton sequence"); }
            else if (!std::strncmp(k, "gradient", 8)) out->use_gradients = 1;
            else if (!std::strncmp(k, "lambda2", 7)) out->use_lambda2 = 1;
            else return fail(VV_ERR_INVALID, std::string("Unrecognized argument: ") + a);
            if (!ok) return fail(VV_ERR_INVALID, "argument too long");
        } else if (len == 2 && a[0] == '-') {                                          // short form, :118-253
            char *dst = nullptr;
            const char *what = "";
            switch (a[1]) {
            case 'h': out->show_help = 1; break;
            case 'g': out->use_gradients = 1; break;
            case 'l': out->use_lambda2 = 1; break;
            case 'f': dst = out->filter_file; what = "filter kernel (png)"; break;
            case 'n': dst = out->noise_file; what = "noise"; break;
            case 't': dst = out->tf_file; what = "transfer function (png)"; break;
            case 'r': dst = out->redirect_file; what = "redirection"; break;
            case 's': dst = out->halton_file; what = "halton sequence"; break;
            default: return fail(VV_ERR_INVALID, std::string("Unknown argument: ") + a);
            }
            if (dst) {
                if (idx + 1 >= argc) return fail(VV_ERR_INVALID, std::string("Missing filename:  ") + what);
                if (argv[idx + 1][0] == '-') return fail(VV_ERR_INVALID, std::string("Invalid argument: ") + a);
                if (!take(dst, argv[idx + 1])) return fail(VV_ERR_INVALID, "argument too long");
                ++idx;
            }
        } else {                                                                       // positional, :255-270
            if (out->vol_file[0]) return fail(VV_ERR_INVALID, "Unknown arguments (too much filenames).");
            if (!take(out->vol_file, a)) return fail(VV_ERR_INVALID, "argument too long");
        }
    }
    return VV_OK;
}

} // extern "C"
