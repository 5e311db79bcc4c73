// Stand-in for <pcl/filters/voxel_grid.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../../ssf_mini_pcl.h"
