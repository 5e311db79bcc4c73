// Stand-in for <pcl/features/normal_3d.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../../ssf_mini_pcl.h"
