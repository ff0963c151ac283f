// TEST INFRASTRUCTURE ONLY -- headless shim for <drawstuff/drawstuff.h> (ODE's demo viewer; visualization.h:14,
// geom.cpp:2), absent from /root/reference and from this image.  Drawing calls do nothing.  dsSimulationLoop()
// (visualization.cpp:197-205) runs the reference's own step callback without a window: start(), then step(0)
// until ds_shim_max_steps iterations have run (oracle/shim/ode_world.cpp) -- this is how the fall sweep drives
// modelplayer::simulate_ode (player.cpp:326-340) on the CPU.
#ifndef ORACLE_SHIM_DRAWSTUFF_H
#define ORACLE_SHIM_DRAWSTUFF_H
#define DS_VERSION 0x0002
typedef struct dsFunctions {
  int version;
  void (*start)();
  void (*step)(int pause);
  void (*command)(int cmd);
  void (*stop)();
  const char* path_to_textures;
} dsFunctions;
extern long ds_shim_max_steps;   // iterations dsSimulationLoop runs (default 0: return at once)
extern long ds_shim_steps_done;
void dsSimulationLoop(int argc, char** argv, int window_width, int window_height, dsFunctions* fn);
inline void dsSetViewpoint(float*, float*) {}
inline void dsSetColor(float, float, float) {}
inline void dsSetColorAlpha(float, float, float, float) {}
inline void dsDrawSphere(const double*, const double*, double) {}
inline void dsDrawBox(const double*, const double*, const double*) {}
inline void dsDrawCapsule(const double*, const double*, double, double) {}
inline void dsDrawCylinder(const double*, const double*, double, double) {}
inline void dsDrawLine(const double*, const double*) {}
inline void dsDrawTriangle(const double*, const double*, const double*, const double*, const double*, int) {}
#endif
