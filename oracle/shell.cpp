/* oracle/shell.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).  Shell groups (placeholder until
 * the QEPH / BT restatements land). */
#include "oracle.h"
#include <cstdlib>
struct OrcShellGroup { int nel=0; };
void orc_shell_group_free(OrcShellGroup* g){ delete g; }
void orc_shell_dispatch(Oracle&, OrcShellGroup&, double&, int&, int&){ abort(); }
