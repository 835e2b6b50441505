#include "../vv_gl_stub.h"
