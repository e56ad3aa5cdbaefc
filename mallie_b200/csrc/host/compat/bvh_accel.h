// bvh_accel.h -- forwarding header: code written against the reference's bvh_accel.h builds against mallie_b200.
#include "../mallie_api.h"
