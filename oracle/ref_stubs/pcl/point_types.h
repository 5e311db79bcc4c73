// Stand-in for <pcl/point_types.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../ssf_mini_pcl.h"
