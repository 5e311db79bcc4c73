// Stand-in for <pcl/io/pcd_io.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../../ssf_mini_pcl.h"
