/* oracle/shell_bt.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Belytschko-Tsay shell CFORC3
 * (placeholder until the restatement lands). */
#include "shell.h"
#include <cstdlib>
void orc_cforc3(Oracle&, OrcShellGroup&, double&, int&, int&){ abort(); }
