// material.h -- forwarding header: code written against the reference's material.h builds against mallie_b200.
#include "../mallie_api.h"
