// TEST INFRASTRUCTURE ONLY -- link stubs for effdata.cpp, the one reference source that cannot be compiled here:
// it includes "balltreelib1/balltree.h", which is not in the tree (effdata.cpp:4; only the binary btlib/libballtree.a
// is), and reads the ball tree's private layout (effdata.cpp:53).  cpc.cpp calls these entry points when the CPC
// controller runs (OUT OF SCOPE); the gait-evaluation path never reaches them.
#include <cstdio>
#include <cstdlib>
#include "effdata.h"

static void out_of_scope(const char* what) {
  std::fprintf(stderr, "oracle/_ref: %s needs the binary-only ball tree (effdata.cpp), which is out of scope\n", what);
  std::abort();
}
void efficientdata::prepare_data(const list<vector<double> >&, int) { out_of_scope("efficientdata::prepare_data"); }
void efficientdata::prepare_data(double**, int, int) { out_of_scope("efficientdata::prepare_data"); }
int efficientdata::get_dim() { out_of_scope("efficientdata::get_dim"); return 0; }
void efficientdata::get_gammat0s(double*, double, double, map<int, double>&, map<int, double>&) { out_of_scope("efficientdata::get_gammat0s"); }
void efficientdata::get_tpis(list<int>&, int, double*, double, double, double, double) { out_of_scope("efficientdata::get_tpis"); }
