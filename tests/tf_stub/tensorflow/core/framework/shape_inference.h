#pragma once
#include "tensorflow/core/framework/op.h"
