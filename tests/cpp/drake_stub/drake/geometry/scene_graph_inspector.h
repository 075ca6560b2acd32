// stand-in for <drake/geometry/scene_graph_inspector.h>: see ../stub_impl.h
#pragma once
#include "drake/stub_impl.h"
