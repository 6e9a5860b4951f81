#include "common.cuh"
#include "internal.h"
namespace pifu {
struct OctreeState {};
void octree_free(OctreeState* s) { delete s; }
}
