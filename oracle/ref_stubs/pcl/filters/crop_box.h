// Stand-in for <pcl/filters/crop_box.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../../ssf_mini_pcl.h"
