// Included by the reference's solver-facing headers (Include/Simulator.h:20-21, Include/QtSPHSolver.h:21) but nothing
// of it is used on the solver path: intentionally empty.
#pragma once
