// shim: QFunction sources that include <ceed.h> only need the scalar/index types at JIT time
#include <ceed/types.h>
