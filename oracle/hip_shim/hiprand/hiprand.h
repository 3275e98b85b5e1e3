// TEST INFRASTRUCTURE ONLY: the reference includes <hiprand/hiprand.h> but never
// calls a hiprand symbol on this path, so this is empty.
#pragma once
