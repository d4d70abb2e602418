// TEST INFRASTRUCTURE ONLY -- empty stand-in for cereal v1.2.2 (not vendored).
// The reference's serialize() members are templates and are only instantiated
// by oracle/ref_driver.cpp's own in-memory archive.
#pragma once
