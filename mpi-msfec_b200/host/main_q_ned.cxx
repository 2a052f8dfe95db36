// MsFEC_Q_Ned executable (reference: source/main_q_ned.cxx): "-p parameter_file.prm".
#include "basis.h"
int main(int argc, char **argv) { return msfec::driver_main(argc, argv, MSFEC_Q_NED, "Q_Ned"); }
