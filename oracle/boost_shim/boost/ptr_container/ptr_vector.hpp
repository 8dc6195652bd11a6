// empty stand-in (oracle build only)
#pragma once
