#pragma once
#define DS2I_SOURCE_DIR "/root/reference"
