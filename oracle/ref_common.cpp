/* ref_common.cpp -- TEST INFRASTRUCTURE ONLY: gl_* state of the GLSL shim */
#include "glsl_shim.h"
#include "ref_api.h"

namespace glsl {
thread_local uint32_t g_fetch_count[16];
thread_local float g_fbo_dest[4];
thread_local vec4 gl_TexCoord[8];
thread_local vec4 gl_FragColor;
thread_local vec4 gl_FragCoord;
mat4 gl_ModelViewMatrixInverse;
mat4 gl_ModelViewProjectionMatrix;
LightSource gl_LightSource[1];
}

extern "C" void vvref_set_gl_state(const RefUniforms *u)
{
    using namespace glsl;
    gl_ModelViewMatrixInverse[3] = vec4(u->camera[0], u->camera[1], u->camera[2], u->camera[3]);
    gl_LightSource[0].position = vec4(u->light_position[0], u->light_position[1], u->light_position[2], u->light_position[3]);
    gl_LightSource[0].ambient = vec4(u->light_ambient[0], u->light_ambient[1], u->light_ambient[2], u->light_ambient[3]);
    gl_LightSource[0].diffuse = vec4(u->light_diffuse[0], u->light_diffuse[1], u->light_diffuse[2], u->light_diffuse[3]);
    gl_LightSource[0].specular = vec4(u->light_specular[0], u->light_specular[1], u->light_specular[2], u->light_specular[3]);
    gl_LightSource[0].spotExponent = u->spot_exponent;
}
