// stand-in (oracle/ref_stubs_filter): everything ExponentialFilter needs is in filter_stubs.h
#pragma once
#include "deal.II/filter_stubs.h"
