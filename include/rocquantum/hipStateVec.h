/* Reference include path (rocquantum/include/rocquantum/hipStateVec.h): client code that says
 * #include "rocquantum/hipStateVec.h" compiles unchanged against this tree. */
#include "../hipStateVec.h"
