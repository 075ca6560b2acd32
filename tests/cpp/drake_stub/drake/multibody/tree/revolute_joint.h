// stand-in for <drake/multibody/tree/revolute_joint.h>: see ../stub_impl.h
#pragma once
#include "drake/stub_impl.h"
