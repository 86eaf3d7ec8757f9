#include "tps_stub_decls.hpp"
