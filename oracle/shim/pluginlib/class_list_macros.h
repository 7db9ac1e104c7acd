#pragma once
#define PLUGINLIB_EXPORT_CLASS(a, b)
