// Stand-in for <pcl/kdtree/kdtree_flann.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../../ssf_mini_pcl.h"
