// Forwarding header: consumers that include <KokkosSparse_CrsMatrix.hpp> (solver.hpp:7-9, mainwindow.cpp) get the host stand-ins.
#pragma once
#include "b200_dropin.hpp"
