// Drop-in `sub_sampler` executable: same command line as the reference's
// (SubSampler.cpp:667-803); the work happens in libspsp_host / libspsp_b200.
#include "spsp_host.h"
int main(int argc, char **argv) { return spsph_sub_sampler_main(argc, argv); }
