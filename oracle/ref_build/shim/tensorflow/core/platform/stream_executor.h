// Empty stand-in (the CPU plan never uses StreamExecutor). Test infrastructure only.
#pragma once
