// TEST INFRASTRUCTURE ONLY -- shim for ODE's demo header "texturepath.h" (visualization.h:15); see ode/ode.h.
#ifndef ORACLE_SHIM_TEXTUREPATH_H
#define ORACLE_SHIM_TEXTUREPATH_H
#define DRAWSTUFF_TEXTURE_PATH "."
#endif
