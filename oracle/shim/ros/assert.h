#pragma once
#include <ros/console.h>
