// Forwarding header: consumers that include <Kokkos_Core.hpp> (solver.hpp:7-9, mainwindow.cpp) get the host stand-ins.
#pragma once
#include "b200_dropin.hpp"
