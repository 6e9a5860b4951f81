#include "common.cuh"
#include "internal.h"
namespace pifu {
struct McState {};
void mc_free(McState* s) { delete s; }
}
