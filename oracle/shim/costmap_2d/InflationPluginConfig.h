#pragma once
// defaults of cfg/InflationPlugin.cfg
namespace costmap_2d { struct InflationPluginConfig { bool enabled = true; double cost_scaling_factor = 10.0; double inflation_radius = 0.55; }; }
