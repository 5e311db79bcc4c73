// Stand-in for <pcl_conversions/pcl_conversions.h>: nothing of it is used by the files compiled into oracle/_ref.
