/* same include path as the reference's include/piplib/piplib_dp.h:31-47 */
#ifndef PIPLIB_B200_PIPLIB_DP_H
#define PIPLIB_B200_PIPLIB_DP_H
#undef PIPLIB_INT_SP
#undef PIPLIB_INT_GMP
#ifndef PIPLIB_INT_DP
#define PIPLIB_INT_DP 1
#endif
#undef LINEAR_VALUE_IS_LONG
#undef LINEAR_VALUE_IS_MP
#undef LINEAR_VALUE_IS_LONGLONG
#define LINEAR_VALUE_IS_LONGLONG 1
#include <piplib/piplib.h>
#endif
