// TBB stand-in (test infrastructure): Source/Simulator.cpp:38-39 creates a task_scheduler_init; the B200 solver has no
// TBB arena (CUDA grid launches replace the parallel_for loops).
#pragma once
namespace tbb { class task_scheduler_init { public: static const int automatic = -1; explicit task_scheduler_init(int = automatic) {} }; }
