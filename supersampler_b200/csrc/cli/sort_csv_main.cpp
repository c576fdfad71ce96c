#include "spsp_host.h"
int main(int argc, char **argv) { return spsph_sort_csv_main(argc, argv); }
