// stand-in for <drake/geometry/shape_specification.h>: see ../stub_impl.h
#pragma once
#include "drake/stub_impl.h"
