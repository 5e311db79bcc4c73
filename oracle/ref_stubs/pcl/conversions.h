// Stand-in for <pcl/conversions.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../ssf_mini_pcl.h"
