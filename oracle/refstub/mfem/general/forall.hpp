#pragma once
#include "../../mfem.hpp"
