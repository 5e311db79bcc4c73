// Stand-in for <pcl/filters/extract_indices.h>: see ssf_mini_pcl.h (test infrastructure, not PCL).
#include "../../ssf_mini_pcl.h"
