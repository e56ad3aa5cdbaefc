// importers/mesh_loader.h -- forwarding header (MeshLoader::LoadObj / LoadESON).
#include "../../mallie_api.h"
