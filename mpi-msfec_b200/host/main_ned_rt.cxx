// MsFEC_Ned_RT executable (reference: source/main_ned_rt.cxx): "-p parameter_file.prm".
#include "basis.h"
int main(int argc, char **argv) { return msfec::driver_main(argc, argv, MSFEC_NED_RT, "Ned_RT"); }
