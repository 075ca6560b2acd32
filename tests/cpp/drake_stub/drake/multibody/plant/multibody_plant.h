// stand-in for <drake/multibody/plant/multibody_plant.h>: see ../stub_impl.h
#pragma once
#include "drake/stub_impl.h"
