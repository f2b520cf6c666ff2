#pragma once
namespace Banana { class ParticleSystemData {}; }
