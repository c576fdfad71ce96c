// Drop-in `sub_sampler` executable: same command line as the reference's
// (SubSampler.cpp:667-803); the work happens in libspsp_host / libspsp_b200.
#include "spsp_host.h"
#include <cstdio>
#include <unistd.h>
int main(int argc, char **argv)
{
    const int rc = spsph_sub_sampler_main(argc, argv);
    // every output file is closed by now: leave without tearing the CUDA context down (0.2-0.5 s in a short run)
    fflush(stdout);
    fflush(stderr);
    _exit(rc);
}
