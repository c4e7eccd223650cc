"""Case definitions (bridge, damped_oscillator, simple_beam) -- import the submodule you need:
`from dpivae_b200.cases import bridge; bridge.definition, bridge.presets`."""
