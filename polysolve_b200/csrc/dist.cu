// Multi-GPU plumbing (row-range partition, halo exchange, dot all-reduce). Placeholder entry points.
#include "../../include/psb200.h"
#include <cstring>
extern "C" {
int psb200_dist_unique_id(char id128[128]) { std::memset(id128, 0, 128); return PSB200_ERR_COMM; }
int psb200_dist_init(psb200_handle, int, int, const char *) { return PSB200_ERR_COMM; }
}
