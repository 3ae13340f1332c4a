/* same include path as the reference's include/piplib/piplib64.h:32 */
#ifndef PIPLIB_B200_PIPLIB64_H
#define PIPLIB_B200_PIPLIB64_H
#include <piplib/piplib_dp.h>
#endif
