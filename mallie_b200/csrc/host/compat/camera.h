// camera.h -- forwarding header: code written against the reference's camera.h builds against mallie_b200.
#include "../mallie_api.h"
