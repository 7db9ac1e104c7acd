#pragma once
#include <boost/thread.hpp>
