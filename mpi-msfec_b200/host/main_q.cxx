// MsFEC_Q executable (reference: source/main_q.cxx): "-p parameter_file.prm".
#include "basis.h"
int main(int argc, char **argv) { return msfec::driver_main(argc, argv, MSFEC_Q, "Q"); }
