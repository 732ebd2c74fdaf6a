// Helpers shared by the translation units of libungar_b200.so (not part of the public ABI).
#pragma once

int ub_set_error(int code, const char* message);  // records the thread's last error (ungar_b200_last_error) and returns `code`
void ub_count_launch();                           // ungar_b200_launch_count accounting
