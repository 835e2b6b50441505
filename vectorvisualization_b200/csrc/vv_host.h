// vv_host.h -- host-side pieces shared between the C-ABI (vv_renderer.cu) and the loaders (vv_io.cpp)
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/vv_c_api.h"

namespace vvb200 {

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);
const std::string &last_error_string();

// PNG (8-bit gray / gray+alpha / RGB / RGBA / palette, non-interlaced) via zlib; rows top to bottom
bool png_read_file(const char *path, std::vector<uint8_t> &data, int &w, int &h, int &channels, std::string &err);
bool png_write_file(const char *path, const uint8_t *data, int w, int h, int channels, std::string &err);

// DatFile (VV/reader.cpp:81-305)
int parse_dat(const char *path, VVDatInfo *out);
int read_raw(const VVDatInfo *info, int time_step, void *out, size_t out_bytes);
size_t dat_bytes(const VVDatInfo *info);

// NoiseDataSet::loadRawData (VV/dataset.cpp:1347-1389)
int read_noise_file(const char *path, std::vector<uint8_t> &data, int dims[3]);

// TransferEdit::setTFFileNames / loadRGBATF / loadAlphaOpacTF (VV/transferEdit.cpp:113-337): updates tf[256*5]
int load_tf_png(const char *name, uint8_t *tf);

int next_pow2(int v);

} // namespace vvb200
