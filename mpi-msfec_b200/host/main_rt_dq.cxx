// MsFEC_RT_DQ executable (reference: source/main_rt_dq.cxx): "-p parameter_file.prm".
#include "basis.h"
int main(int argc, char **argv) { return msfec::driver_main(argc, argv, MSFEC_RT_DQ, "RT_DQ"); }
