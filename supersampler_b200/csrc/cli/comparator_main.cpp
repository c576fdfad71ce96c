// Drop-in `comparator` executable: same command line as the reference's
// (Comparator.cpp:464-521).
#include "spsp_host.h"
int main(int argc, char **argv) { return spsph_comparator_main(argc, argv); }
