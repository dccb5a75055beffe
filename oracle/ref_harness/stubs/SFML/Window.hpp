// Headless stand-in for <SFML/Window.hpp>: everything the harness needs is in Graphics.hpp.
#pragma once
