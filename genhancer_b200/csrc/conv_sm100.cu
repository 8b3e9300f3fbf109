#include "internal.h"
namespace gh { int conv_init() { return GH_OK; } }
