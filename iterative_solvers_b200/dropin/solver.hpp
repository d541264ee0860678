// Forwarding header: keeps the reference's include name (solver/solver.hpp); the classes live in b200_dropin.hpp.
#pragma once
#include "b200_dropin.hpp"
