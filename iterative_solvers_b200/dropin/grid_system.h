// Forwarding header: keeps the reference's include name (solver/grid_system.h); the classes live in b200_dropin.hpp.
#pragma once
#include "b200_dropin.hpp"
