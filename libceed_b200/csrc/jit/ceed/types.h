// Minimal stand-in for <ceed/types.h> so that user QFunction headers JIT-compile when the b200 C ABI is used without a
// libCEED installation.  Only what QFunction sources may reference is provided; when the backend runs under libCEED the
// reference's own header is found first on the include path.
#ifndef CEED_B200_TYPES_SHIM_H
#define CEED_B200_TYPES_SHIM_H
typedef int       CeedInt;
typedef long long CeedSize;
typedef signed char CeedInt8;
typedef double    CeedScalar;
#define CEED_EPSILON 1e-16
#define CeedInt_FMT "d"
#ifndef CEED_QFUNCTION
#define CEED_QFUNCTION(name) static int name
#endif
#ifndef CEED_QFUNCTION_HELPER
#define CEED_QFUNCTION_HELPER static inline
#endif
#ifndef CEED_Q_VLA
#define CEED_Q_VLA Q
#endif
#ifndef CeedPragmaSIMD
#define CeedPragmaSIMD
#endif
#define CeedIntMin(a, b) ((a) < (b) ? (a) : (b))
#define CeedIntMax(a, b) ((a) > (b) ? (a) : (b))
typedef enum { CEED_ERROR_SUCCESS = 0, CEED_ERROR_MINOR = 1, CEED_ERROR_MAJOR = -1, CEED_ERROR_BACKEND = -2, CEED_ERROR_UNSUPPORTED = -3 } CeedErrorType;
#endif
