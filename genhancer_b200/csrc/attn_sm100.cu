#include "internal.h"
namespace gh { int attn_init() { return GH_OK; } }
